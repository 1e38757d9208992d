// Shared host/device helpers for libsag.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include <map>

#include "../../include/sag.h"

namespace sag {

void set_error(const char* fmt, ...);
extern thread_local int g_launch_count;   // kernels launched by this thread since last reset

#define SAG_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      sag::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),    \
                     cudaGetErrorString(_e));                                                  \
      return SAG_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

#define SAG_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      sag::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

#define SAG_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != SAG_OK) return _r; \
  } while (0)

#define SAG_LAUNCH_CHECK()                       \
  do {                                           \
    sag::g_launch_count++;                       \
    SAG_CHECK_CUDA(cudaGetLastError());          \
  } while (0)

// ---- optional per-launch CUDA-event profiling (sag_set_option "profile"): bench.py's roofline numbers ----
enum ProfCat { PROF_CONV = 0, PROF_DECONV, PROF_FC, PROF_STFT, PROF_ISTFT, PROF_POINTWISE, PROF_MIX, PROF_NCAT };
// flops: useful work (every product of the reference graph that can reach the output, once); issued: what the kernel
// multiplies (sub-pixel transposed convs carry zero taps / border cells, conv1's space-to-depth form pads K 147 -> 256)
struct ProfRec { int cat; double flops; double bytes; cudaEvent_t e0, e1; double issued; char name[48]; int tile, split; };
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  void clear();
};
extern thread_local Profiler* g_prof;   // set by forward() while profiling is enabled
struct ProfScope {                      // records an event pair around the launches issued in its lifetime
  cudaStream_t st;
  int idx = -1;
  ProfScope(int cat, double flops, double bytes, cudaStream_t s, const char* name = nullptr, double issued = -1.0, int tile = 0,
            int split = 0);
  ~ProfScope();
};

// ---- programmatic dependent launch: the kernels of the forward are launched with programmatic stream serialisation,
// so the next kernel's launch and prologue overlap the tail of the current one.  Every such kernel calls
// pdl_prologue() first: it lets its own dependent start launching and then waits until the kernel before it in the
// stream has completed and flushed its writes (transitively: everything earlier in the stream).  SAG_PDL=0 disables it.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at = {};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  if (pdl_enabled()) { cfg.attrs = &at; cfg.numAttrs = 1; }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// Gather-GEMM geometry shared by conv / transposed-conv phases / FC.
//   rows   m = ((n*PH + i)*PW + j)                        output positions of this launch
//   out pixel (oy0 + i*osy, ox0 + j*osx); address y + n*y_sn + oy*y_sh + ox*y_sw + co*y_sc
//   tap t reads input pixel (i*isy + dy[t], j*isx + dx[t]) (zero outside [0,H)x[0,W)), channels ci,
//   weight slab w + widx[t]*Cin*Cout laid out [Cin][Cout]
// ------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 112;

struct GatherGeom {
  int N, H, W, Cin;       // input tensor
  int64_t x_ld;           // input pixel stride (elements); input channel offset folded into pointer
  int64_t x_row;          // elements between image rows (0: W * x_ld); tcgen05 path only -- lets pixels overlap (x_ld < Cin)
  int PH, PW;             // iteration grid
  int isy, isx;           // input step per grid step
  int oy0, ox0, osy, osx; // output placement
  int64_t y_sn, y_sh, y_sw, y_sc;  // output strides (elements)
  int Cout;
  int T;                  // taps
  short dy[kMaxTaps], dx[kMaxTaps], widx[kMaxTaps];
};

struct Epilogue {
  const float* bias;      // [Cout] or null
  int relu;
  // batch-norm statistics (sum, sum of squares over the rows) or null: [Cout][2] fixed-point accumulators (fx_atomic_add),
  // zeroed by the caller, accumulated with integer atomics -- exact, hence independent of the arrival order of the CTAs
  unsigned long long* stat_sum;
  unsigned long long* stat_sqs;
  // mask-gain fusion (deconv1 of the U-Net decoder, model.py:326-334 + 424-432 by linearity): instead of the 32 track logits
  // of a time-frequency bin the epilogue writes G[o, seg] = sum_k W[o, k, seg] * sigmoid(logit_k) for the 3 x 3 (channel,
  // localization segment) pairs: gains (rows, 9, frames, bins) from gain_loc (rows, 3 segments, 3*(tracks+1)); null = off
  const float* gain_loc;
  float* gains;
  int64_t gain_plane;     // frames * bins
  // stream-K (tcgen05 path): UMMA_SK_FLAGS zeroed ints owned by the caller's stream (a kernel leaves them zeroed); null = the
  // launch never deals K chunks of a tile out to several CTAs
  int* sk_flags;
};
constexpr int UMMA_SK_FLAGS = 1024;

// Exact fixed-point accumulation of float partial sums: word 0 = integer part (two's complement), word 1 = fraction * 2^40.
// A float at or above 2^-17 in magnitude is represented exactly (24-bit mantissa), smaller ones are truncated to the 2^-40 grid;
// integer addition is associative, so the total does not depend on the order in which CTAs arrive.  Up to 2^12 partials
// (fraction word stays below 2^52), partial sums below 2^62.
#ifdef __CUDACC__
__device__ __forceinline__ void fx_atomic_add(unsigned long long* acc2, float p) {
  const double d = (double)p;
  const double fl = floor(d);
  atomicAdd(acc2, (unsigned long long)(long long)fl);
  atomicAdd(acc2 + 1, (unsigned long long)((d - fl) * 1099511627776.0));
}
#endif
#ifdef __CUDACC__
// tf.sigmoid (model.py:334) for the mask path, on the special-function unit: 1 / (1 + 2^(-x * log2 e)) with ex2.approx and
// rcp.approx (relative error ~3e-7, five instructions).  mask_gains_kernel (fft.cu) and the fused deconv1 epilogue
// (conv_umma.cu) both use exactly this sequence, so the two paths produce the same gains bit for bit.
__device__ __forceinline__ float sigmoid_sfu(float x) {
  float e, r;
  const float t = __fmul_rn(x, -1.4426950408889634f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  const float d = __fadd_rn(1.f, e);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}
#endif
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline double fx_value(const unsigned long long* acc2) {
  return (double)(long long)acc2[0] + (double)acc2[1] * (1.0 / 1099511627776.0);
}

// TF 'SAME' padding (asymmetric): returns pad_before; out = ceil(in/s)
static inline int same_pad_before(int in, int k, int s, int* out) {
  int o = (in + s - 1) / s;
  int total = (o - 1) * s + k - in;
  if (total < 0) total = 0;
  *out = o;
  return total / 2;
}

// FFMA gather-GEMM (conv_ffma.cu)
int launch_gather_gemm_ffma(const float* x, const float* w, float* y, const GatherGeom& g, const Epilogue& ep,
                            cudaStream_t st);

// ---- tcgen05 path (conv_umma.cu): SAG_PREC_BF16 / SAG_PREC_BF16X3 -----------------------------------------------
// B operand of one layer, packed once: per (N tile, K chunk) a bf16 hi plane (+ lo plane for BF16X3) already in the
// swizzled K-major shared-memory layout of tcgen05.mma, so each pipeline stage fetches it with one bulk-async copy.
struct UmmaWeights {
  void* packed = nullptr;
  int K = 0, KC = 0, N = 0, BN = 0, NT = 0, planes = 0;
  // output column map of the sub-pixel transposed conv (null for conv / FC): element offset, row / column displacement
  // and bias per GEMM column
  int* col_off = nullptr;
  short* col_dy = nullptr;
  short* col_dx = nullptr;
  float* col_bias = nullptr;
  int vec4 = 0;
  int run8 = 0;         // mapped output: every group of 8 columns is 8 consecutive floats of one output row
  int order = 0;        // column order of a sub-pixel transposed conv (decode_subpixel_column in conv_umma.cu)
  int64_t M_hint = 0;   // row count the tile width was chosen for
  int a_single = 0;     // the activation this image multiplies has ONE exact bf16 plane (uint8 frames as 2k - 255): A x (B_hi + B_lo)
  // tiled tensor map of the packed image (rows of 128 bytes, boxes of BN/2 rows) for the CTA-pair kernel, which fetches its
  // weight blocks with .cta_group::2 tensor copies (a CUtensorMap, kept opaque here); wmap_ok == 0: not available
  alignas(64) unsigned char wmap[128] = {0};
  int wmap_ok = 0;
};
void umma_free(UmmaWeights* w);
// tile width of a contraction with N columns over M rows (a single M tile takes narrow tiles: see conv_umma.cu)
int umma_tile_width(int K, int N, int64_t M);
// wk: device fp32 [K][N] with row stride ldw (conv HWIO / FC [in,out] are already in this form); M: rows of the GEMM
// the image will be used for (selects the tile width)
int umma_pack_weights(const float* wk, int K, int N, int64_t ldw, int precision, int64_t M, UmmaWeights* out, cudaStream_t st);
// stride-2 conv weights HWIO re-expressed for the 2x2 space-to-depth image with 16-channel pixels (K = ceil(kh/2)*ceil(kw/2)*16)
int umma_pack_conv_s2d(const float* w_hwio, int kh, int kw, int cin, int cout, int precision, int64_t M, int int_frames, UmmaWeights* out,
                       cudaStream_t st);
// conv1 on uint8 video frames as one exact bf16 plane of integers: true when the halo kernel takes the layer at this size
bool umma_int_frames_supported(int n, int oh, int ow);
// w_hwoi: device tf.nn.conv2d_transpose weights [kh,kw,Cout,Cin]; order 0: columns (py,px,co), 1: columns (py,co,px)
int umma_pack_deconv(const float* w_hwoi, const float* bias, int kh, int kw, int cout, int cin, int sh, int sw, int order,
                     int64_t y_sh, int64_t y_sw, int64_t y_sc, int precision, int64_t M, UmmaWeights* out, cudaStream_t st);
// Activation view handed to the tensor-core kernels.  ACT_F32: p = float*.  ACT_BF2: the value x is stored as two bf16
// planes, hi = bf16(x) at p and lo = bf16(x - hi) at p + plane bytes (plane == 0: hi only, SAG_PREC_BF16); strides of
// the geometry are in elements either way.
enum { ACT_F32 = 0, ACT_BF2 = 1 };
struct ActView {
  void* p = nullptr;
  int fmt = ACT_F32;
  int64_t plane = 0;
  ActView() {}
  ActView(const float* f) : p(const_cast<float*>(f)) {}
  ActView(void* q, int f, int64_t pl) : p(q), fmt(f), plane(pl) {}
};
extern thread_local int g_umma_pair;  // -1: SAG_UMMA_PAIR env (default off); 0/1: forced: CTA pairs (cta_group::2) on the TMA path
extern thread_local int g_umma_tma;   // -1: SAG_UMMA_TMA env (default on); 0/1: forced for this thread's launches
extern thread_local int g_umma_halo;  // -1: SAG_UMMA_HALO env (default on); 0/1: forced: halo-resident kernel for the 3x3 stride-1 convs
// scratch: split-K workspace of at least the bytes umma_split_k reports (null: never split)
int umma_split_k(int K, int N, int64_t M, size_t* scratch_bytes);
// stream-K pieces of one cluster in execution order (items: [n][5] = tile, first chunk, end chunk, produce, fix_first); returns n
bool umma_stream_k(int K, int N, int64_t M);      // the planner deals this shape out as stream-K (given flags + scratch)
int streamk_schedule(int64_t tiles, int kc, int clusters, int cluster, int* items, int max_items);
int launch_gather_gemm_umma(const ActView& x, const UmmaWeights& w, const ActView& y, const GatherGeom& g, const Epilogue& ep,
                            int oh_lim, int ow_lim, float* scratch, cudaStream_t st);

// helpers building geometries (geom.cu)
int make_conv_geom(GatherGeom* g, int n, int h, int w, int cin, int64_t x_ld, int kh, int kw, int cout, int sh, int sw,
                   int same_pad, int64_t y_ld, int* oh, int* ow);
// transposed conv (VALID) phase (py,px) restricted to output rows [row0,row1) ; output strides for a tensor whose
// row 0 is full-output row `row0`.  Returns 1 if the phase is empty.
int make_deconv_phase_geom(GatherGeom* g, int n, int h, int w, int cin, int64_t x_ld, int kh, int kw, int cout, int sh,
                           int sw, int py, int px, int row0, int row1, int64_t y_sn, int64_t y_sh, int64_t y_sw,
                           int64_t y_sc);

// sub-pixel GEMM geometry of a whole transposed conv restricted to output rows [row0,row1) (conv_umma.cu)
int make_deconv_subpixel_geom(GatherGeom* g, int n, int h, int w, int cin, int64_t x_ld, int kh, int kw, int sh, int sw,
                              int row0, int row1, int64_t y_sn, int64_t y_sh, int64_t y_sw, int64_t y_sc, int* oh_lim,
                              int* ow_lim);

// pointwise.cu
// Batch statistics of one BN layer as the conv epilogues leave them: the consumer kernels turn them into per-channel
// scale / shift in their prologue (no separate finalize launch).  scale = gamma*rsqrt(var+eps), shift = beta - mean*scale,
// biased variance (contrib batch_norm, core.py:209-210).
struct BnStats {
  const unsigned long long* sum = nullptr;      // [C][2] fixed-point sums (fx_atomic_add / fx_value)
  const unsigned long long* sqs = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  double inv_count = 0.0;
  float eps = 1e-3f;
};
int launch_bn_apply_stats(const float* x, const BnStats& bn, const ActView& residual, int relu, const ActView& y, int64_t rows,
                          int c, cudaStream_t st);
int launch_bn_relu_maxpool_stats(const float* x, const BnStats& bn, int n, int h, int w, int c, const ActView& y, cudaStream_t st);
int launch_bn_relu_maxpool(const float* x, const float* scale, const float* shift, int n, int h, int w, int c,
                           const ActView& y, cudaStream_t st);   // 3x3/2 SAME; scale==null -> plain max-pool
int launch_channel_stats(const float* x, int64_t rows, int c, unsigned long long* sum, unsigned long long* sqs, cudaStream_t st);
int launch_tile_rows(const ActView& src, int64_t src_ld, const ActView& dst, int64_t dst_ld, int groups, int reps, int c,
                     cudaStream_t st);   // dst[(g*reps+r)*dst_ld + :c] = src[g*src_ld + :c]
int launch_mix(const float* x_sep, const float* loc, int batch, int tracks, int t, int segments, float* out,
               cudaStream_t st);
// A batch of input frames (n,h,w,3): prepared fp32, or the uint8 frame as decoded from the jpg (video: x/255 - 0.5,
// myutils.py:88-89; flow: quantised (angle, -, magnitude) + per-frame (min, max) limits, feeder.py:147-161) prepared on the device
// FRAMES_U8_VIDEO_INT: uint8 video frames delivered as the integers 2k - 255 (exact in one bf16 plane; x/255 - 0.5 = (2k - 255)/510,
// the 1/510 lives in the packed conv1 weights: umma_pack_conv_s2d `int_frames`) -- internal to resnet18_tower
enum { FRAMES_F32 = 0, FRAMES_U8_VIDEO = 1, FRAMES_U8_FLOW = 2, FRAMES_U8_VIDEO_INT = 3 };
struct FrameSrc {
  const void* p = nullptr;
  int kind = FRAMES_F32;
  const double* lims = nullptr;      // FRAMES_U8_FLOW: device (n,2) doubles
  FrameSrc() {}
  FrameSrc(const float* f) : p(f) {}
  FrameSrc(const void* q, int k, const double* l) : p(q), kind(k), lims(l) {}
};
// (n,h,w,c<=4) frames -> 2x2 space-to-depth of the image placed at (pt,pl) inside a zero canvas: (n,h2,w2,16) with channel
// (py*2+px)*c + ch = canvas[2*y2+py][2*x2+px][ch], zero beyond 4*c
int launch_space_to_depth16(const FrameSrc& src, int n, int h, int w, int c, int pt, int pl, int h2, int w2, const ActView& out,
                            cudaStream_t st);
int launch_frames_to_f32(const FrameSrc& src, int n, int h, int w, float* out, cudaStream_t st);
int launch_pack_deconv_weights(const float* w_hwoi, float* out, int taps, int cout, int cin, cudaStream_t st);

// fft.cu
int launch_stft(const float* x, int rows, int n_samples, int wind, int hop, int n_frames_total, int frame0,
                int n_frames_out, float* cplx_out, int mag0, int n_mag, const ActView& mag_out, cudaStream_t st);
// masked inverse STFT with overlap-add: S (rows_s, n_frames, wind) complex; mask (rows_s*tracks, n_frames, wind) real
// logits (sigmoid applied inside when apply_sigmoid) or null (mask == 1, tracks == 1);
// out[(row*tracks+k), j] for j in [crop0, crop0+n_out) of the reference istft output.
int launch_istft(const float* S, const float* mask, int apply_sigmoid, int rows_s, int tracks, int n_frames, int wind,
                 int n_overlap, int crop0, int n_out, float* out, cudaStream_t st);

// fused masked inverse STFT + mixing (model.py:334-347 + 424-432 by linearity): out (rows, t_out, 3); see fft.cu
int istft_mix_supported(int tracks, int t_out, int segments, int wind);
size_t istft_mix_gain_floats(int rows, int n_frames, int wind, int segments);
int launch_istft_mix(const float* S, const float* mask, const float* loc, float* gains, int rows, int tracks, int n_frames, int wind,
                     int n_overlap, int crop0, int t_out, int segments, float* out, cudaStream_t st);

// metrics.cu
int launch_metrics(const float* pred, const float* gt, int batch, int t, int audio_rate, float* stft_ps, float* lsd_ps,
                   float* mse_ps, float* snr_ps, float* env_ps, float* amp, void* scratch, cudaStream_t st);
// mel log-spectral distance of myutils.compute_lsd_dist (librosa melspectrogram restated): pred, gt (B,T,3) -> out (B,3)
int launch_mel_lsd(const float* pred, const float* gt, int batch, int t, int audio_rate, float* out, cudaStream_t st);
int launch_sh_rms(const float* ambi, int batch, int t, float ang_res, float* rms, cudaStream_t st);

}  // namespace sag
