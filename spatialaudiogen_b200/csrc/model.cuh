// Handle, weight store, workspace arena and the forward orchestration of SptAudioGen.inference_ops
// (reference model.py:356-434) for libsag.so.
#pragma once
#include "common.cuh"

namespace sag {

struct DevTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  int64_t ld = 0;      // stride of the second-to-last dimension (== shape.back() when contiguous)
  int fmt = 0;         // ACT_F32, or ACT_BF2 (split bf16 planes: p = hi plane, lo plane `plane` bytes further)
  int64_t plane = 0;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// Bump allocator over the caller's workspace.  In dry mode nothing is touched and only the peak is recorded
// (that is how sag_workspace_bytes is computed: by running the same planner).
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  // dry && prepare: additionally build (allocate + pack) the tensor-core weight images this batch size needs -- the only
  // place they are created; the real pass fails if one is missing (no allocation inside sag_forward)
  bool prepare = false;
  bool failed = false;
  // split-K scratch shared by all layers of a pass (they run back to back on one stream): the dry pass records the
  // largest request in scratch_need, the real pass gets that many bytes at `scratch`
  size_t scratch_need = 0, scratch_cap = 0;
  float* scratch = nullptr;
  template <class T>
  T* alloc(int64_t n) {
    off = (off + 255) & ~(size_t)255;
    size_t bytes = (size_t)n * sizeof(T);
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    if (off > peak) peak = off;
    if (!dry && off > cap) failed = true;
    return p;
  }
};

// A tensor of the forward: NHWC pixels with `ld` elements between consecutive pixels, stored as fp32 (FFMA path,
// raw conv outputs awaiting batch-norm, final outputs) or as split-bf16 planes (everything the tensor cores read).
struct Act {
  ActView v;
  int64_t ld = 0;
  Act() {}
  Act(const float* p, int64_t ld_) : v(p), ld(ld_) {}
  float* f32() const { return reinterpret_cast<float*>(v.p); }
  // same tensor, channel range starting at `c`
  Act channels(int64_t c) const {
    Act a = *this;
    a.v.p = reinterpret_cast<char*>(v.p) + c * (v.fmt == ACT_BF2 ? 2 : 4);
    return a;
  }
};

}  // namespace sag

struct sag_handle {
  sag_config cfg;
  sag_dims dims;
  int device = 0;
  std::vector<std::pair<std::string, std::vector<int64_t>>> expected;   // checkpoint layout
  std::map<std::string, sag::DevTensor> weights;                        // TF layout on device
  std::map<std::string, sag::DevTensor> packed;                         // kernel layouts derived at load time
  std::map<std::string, sag::UmmaWeights> umma;                         // tcgen05 operand images, key "<scope>#<precision>"
  std::map<std::string, sag::DevTensor> ends;                           // taps of the last forward
  std::vector<std::string> end_order;
  int last_launches = 0;
  int finalized = 0;
  sag::Profiler prof;    // per-launch event timing of the last forward (option "profile")
  int keep_sep_channels = 0;   // 1: inverse STFT and mixing as two kernels so that `separation/all_channels` (x_sep) exists
  int cta_pair = -1;     // tcgen05 path on CTA pairs (cta_group::2): -1 default (off), 0 / 1 forced
  int int_frames = 1;    // uint8 video frames reach conv1 as one exact bf16 plane of integers (0: as x/255 - 0.5 split in two planes, bit-identical to float frames)
  int halo_conv = -1;    // 3x3 / stride-1 convs on the halo-resident kernel: -1 default (on where eligible), 0 / 1 forced
  int tma_gather = -1;   // activation gather of the tcgen05 path: -1 default (TMA im2col where it applies), 0 cp.async, 1 TMA
  int skip_unused = 1;   // skip mask rows / frames that cannot reach the cropped output (bit-identical result)
  // Branches of the forward that do not depend on each other run on a second stream of the handle, forked from and joined
  // back into the caller's stream with events (still fully asynchronous to the host; captured by CUDA graphs as a fork /
  // join): STFT + audio encoder + audio-fc beside the visual towers, the localization FCs beside the U-Net decoder.  Those
  // are chains of small grids that leave most SMs idle on their own and fill the tails of the towers' persistent kernels.
  int overlap = 1;
  int fuse_gains = 1;    // fold sigmoid + the 32 -> 9 localization-weighted sums into deconv1's epilogue where the tile plan allows
  cudaStream_t side = nullptr;
  cudaStream_t side2 = nullptr;   // the 1x1 / stride-2 shortcut convolutions of the ResNet blocks, beside conv_1 / conv_2 of their block
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* sk_flags = nullptr;        // stream-K flags (Epilogue::sk_flags) of the contractions on the caller's stream
};

namespace sag {

int build_expected(sag_handle* h);
int derive_dims(const sag_config& c, sag_dims* d);
int forward(sag_handle* h, const float* audio, const FrameSrc& video, const FrameSrc& flow, float* out, Arena& ar, int B,
            cudaStream_t st);
int resnet18_tower(sag_handle* h, const std::string& scope, const FrameSrc& x, int B, int H, int W, Act* y, Arena& ar,
                   cudaStream_t st);
int launch_act_to_f32(const ActView& src, float* dst, int64_t n, cudaStream_t st);   // pointwise.cu
const char* last_error_cstr();
int fft_prepare(int n);
void istft_needed_frames(int n_frames, int wind, int n_overlap, int crop0, int n_out, int* f_lo, int* f_hi);
int sh_mesh_dims(float ang_res, int* n_nu, int* n_phi);

// one-shot contraction with unpacked weights (stage entry points): FFMA for SAG_PREC_FP32, else packs and runs tcgen05
int launch_gather_gemm(int precision, const float* x, const float* w, float* y, const GatherGeom& g, const Epilogue& ep,
                       cudaStream_t st);

}  // namespace sag
