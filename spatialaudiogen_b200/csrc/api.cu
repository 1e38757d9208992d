// extern "C" surface of libsag.so (include/sag.h).  Every entry point validates its arguments, returns a
// SAG_E* code and never throws / aborts; see sag.h for the reference interface each one replaces.
#include "model.cuh"

using namespace sag;

namespace {

cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int64_t numel(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto v : s) n *= v;
  return n;
}

bool is_deconv_weight(const std::string& n) {
  return n.rfind("separation/deconv", 0) == 0 && n.size() > 8 && n.compare(n.size() - 8, 8, "/weights") == 0;
}

bool valid_precision(int p) { return p == SAG_PREC_FP32 || p == SAG_PREC_BF16 || p == SAG_PREC_BF16X3 || p == SAG_PREC_MIXED; }

void free_tensor(DevTensor& t) {
  if (t.p) cudaFree(t.p);
  t.p = nullptr;
}

}  // namespace

extern "C" {

const char* sag_last_error(void) { return last_error_cstr(); }
const char* sag_version(void) { return "spatialaudiogen_b200 libsag 0.1 (sm_100a)"; }

// CRC-32C (Castagnoli) of host bytes, continuing from `crc` (0 to start): the payload checksums of TensorFlow V2 checkpoint
// bundles (tf_checkpoint.read_bundle verifies every tensor it restores; deploy.py:79-87).  Slicing by 8.
uint32_t sag_crc32c(const void* host_data, size_t size, uint32_t crc) {
  static uint32_t table[8][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 255];
    init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(host_data);
  crc = ~crc;
  while (size >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= crc;
    crc = table[7][v & 255] ^ table[6][(v >> 8) & 255] ^ table[5][(v >> 16) & 255] ^ table[4][(v >> 24) & 255] ^
          table[3][(v >> 32) & 255] ^ table[2][(v >> 40) & 255] ^ table[1][(v >> 48) & 255] ^ table[0][v >> 56];
    p += 8;
    size -= 8;
  }
  while (size--) crc = table[0][(crc ^ *p++) & 255] ^ (crc >> 8);
  return ~crc;
}

int sag_config_default(sag_config* cfg) {
  SAG_REQUIRE(cfg != nullptr, SAG_EINVAL, "sag_config_default: NULL");
  memset(cfg, 0, sizeof(*cfg));
  cfg->ambi_order = 1;
  cfg->audio_rate = 48000;
  cfg->video_rate = 10;
  cfg->context = 1.0;
  cfg->sample_duration = 0.1;
  cfg->enc_audio = 1;
  cfg->enc_video = 1;
  cfg->enc_flow = 1;
  cfg->separation = SAG_SEP_NONE;              // model.py:31 default 'none'
  cfg->sep_num_tracks = 32;                    // definitions.py:13
  cfg->n_loc_fc = 2;
  cfg->loc_fc_units[0] = 512;                  // definitions.py:16
  cfg->loc_fc_units[1] = 512;
  cfg->sep_fft_window = 0.025;                 // definitions.py:17
  cfg->precision = SAG_PREC_FP32;
  cfg->frame_h = 224;
  cfg->frame_w = 448;
  return SAG_OK;
}

int sag_create(sag_handle** out, const sag_config* cfg) {
  SAG_REQUIRE(out != nullptr && cfg != nullptr, SAG_EINVAL, "sag_create: NULL argument");
  *out = nullptr;
  SAG_REQUIRE(cfg->n_loc_fc >= 0 && cfg->n_loc_fc <= 4, SAG_EINVAL, "sag_create: n_loc_fc %d outside [0,4]", cfg->n_loc_fc);
  SAG_REQUIRE(cfg->separation == SAG_SEP_NONE || cfg->separation == SAG_SEP_UNET_MASK, SAG_EINVAL,
              "Unknown separation mode.");                                            // model.py:351
  SAG_REQUIRE(valid_precision(cfg->precision), SAG_EINVAL, "sag_create: unknown precision %d", cfg->precision);
  SAG_REQUIRE(cfg->sep_num_tracks >= 1 && cfg->sep_num_tracks <= 256, SAG_EINVAL, "sag_create: sep_num_tracks %d", cfg->sep_num_tracks);
  SAG_REQUIRE(cfg->frame_h > 0 && cfg->frame_w > 0, SAG_EINVAL, "sag_create: bad frame size");
  int ndev = 0;
  SAG_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  SAG_REQUIRE(ndev > 0, SAG_ECUDA, "sag_create: no CUDA device (this library has no CPU path)");
  sag_handle* h = new (std::nothrow) sag_handle();
  SAG_REQUIRE(h != nullptr, SAG_ENOMEM, "sag_create: out of host memory");
  h->cfg = *cfg;
  int r = derive_dims(h->cfg, &h->dims);
  if (r == SAG_OK) r = build_expected(h);
  if (r == SAG_OK && cudaGetDevice(&h->device) != cudaSuccess) { set_error("cudaGetDevice failed"); r = SAG_ECUDA; }
  if (r == SAG_OK) r = fft_prepare(h->dims.wind_size);
  if (r == SAG_OK) {
    static const bool overlap_env = [] { const char* v = getenv("SAG_OVERLAP"); return v == nullptr || atoi(v) != 0; }();
    h->overlap = overlap_env ? 1 : 0;
    // highest priority: its short grids take the SMs the persistent tower kernels free at their tails before the next tower
    // kernel's CTAs do, so the side chain advances at every kernel boundary of the main stream
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool ok = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&h->side2, cudaStreamNonBlocking, prio_lo) == cudaSuccess;
    for (int i = 0; i < 6 && ok; ++i) ok = cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming) == cudaSuccess;
    // stream-K flags of the main stream's contractions (zero between launches)
    ok = ok && cudaMalloc(&h->sk_flags, UMMA_SK_FLAGS * sizeof(int)) == cudaSuccess &&
         cudaMemset(h->sk_flags, 0, UMMA_SK_FLAGS * sizeof(int)) == cudaSuccess;
    if (!ok) { set_error("sag_create: could not create the side stream / events"); r = SAG_ECUDA; }
  }
  if (r != SAG_OK) { sag_destroy(h); return r; }
  *out = h;
  return SAG_OK;
}

int sag_destroy(sag_handle* h) {
  if (h == nullptr) return SAG_OK;
  for (auto& kv : h->weights) free_tensor(kv.second);
  for (auto& kv : h->packed) free_tensor(kv.second);
  for (auto& kv : h->umma) umma_free(&kv.second);
  h->prof.clear();
  for (int i = 0; i < 6; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->side2) cudaStreamDestroy(h->side2);
  if (h->sk_flags) cudaFree(h->sk_flags);
  delete h;
  return SAG_OK;
}

int sag_get_dims(const sag_handle* h, sag_dims* out) {
  SAG_REQUIRE(h != nullptr && out != nullptr, SAG_EINVAL, "sag_get_dims: NULL argument");
  *out = h->dims;
  return SAG_OK;
}

int sag_set_option(sag_handle* h, const char* key, int value) {
  SAG_REQUIRE(h != nullptr && key != nullptr, SAG_EINVAL, "sag_set_option: NULL argument");
  std::string k(key);
  if (k == "skip_unused") { h->skip_unused = value ? 1 : 0; return SAG_OK; }
  if (k == "keep_sep_channels") { h->keep_sep_channels = value ? 1 : 0; return SAG_OK; }
  if (k == "cta_pair") { h->cta_pair = value < 0 ? -1 : (value ? 1 : 0); return SAG_OK; }
  if (k == "tma_gather") { h->tma_gather = value < 0 ? -1 : (value ? 1 : 0); return SAG_OK; }
  if (k == "int_frames") { h->int_frames = value ? 1 : 0; return SAG_OK; }
  if (k == "halo_conv") { h->halo_conv = value < 0 ? -1 : (value ? 1 : 0); return SAG_OK; }
  if (k == "profile") { h->prof.on = value != 0; if (!value) h->prof.clear(); return SAG_OK; }
  if (k == "overlap") { h->overlap = value ? 1 : 0; return SAG_OK; }
  if (k == "fuse_gains") { h->fuse_gains = value ? 1 : 0; return SAG_OK; }
  if (k == "precision") {
    SAG_REQUIRE(valid_precision(value), SAG_EINVAL, "unknown precision %d", value);
    h->cfg.precision = value;
    return SAG_OK;
  }
  set_error("sag_set_option: unknown option '%s'", key);
  return SAG_EINVAL;
}

// ---- weights -------------------------------------------------------------------------------------------------
int sag_num_weights_expected(const sag_handle* h) { return h ? (int)h->expected.size() : SAG_EINVAL; }

int sag_weight_name(const sag_handle* h, int i, char* buf, int buflen, int64_t* shape4, int* rank) {
  SAG_REQUIRE(h != nullptr && buf != nullptr && buflen > 0, SAG_EINVAL, "sag_weight_name: bad argument");
  SAG_REQUIRE(i >= 0 && i < (int)h->expected.size(), SAG_EINVAL, "sag_weight_name: index %d out of range", i);
  const auto& e = h->expected[i];
  snprintf(buf, buflen, "%s", e.first.c_str());
  if (rank) *rank = (int)e.second.size();
  if (shape4)
    for (size_t k = 0; k < 4; ++k) shape4[k] = k < e.second.size() ? e.second[k] : 1;
  return SAG_OK;
}

int sag_load_weight(sag_handle* h, const char* tf_name, const float* host_data, const int64_t* shape, int rank) {
  SAG_REQUIRE(h != nullptr && tf_name != nullptr && host_data != nullptr && shape != nullptr, SAG_EINVAL, "sag_load_weight: NULL argument");
  const std::string name(tf_name);
  const std::vector<int64_t>* exp = nullptr;
  for (const auto& e : h->expected)
    if (e.first == name) { exp = &e.second; break; }
  SAG_REQUIRE(exp != nullptr, SAG_EINVAL, "sag_load_weight: variable '%s' is not part of this model", tf_name);
  SAG_REQUIRE(rank == (int)exp->size(), SAG_EINVAL, "sag_load_weight: '%s' has rank %d, expected %d", tf_name, rank, (int)exp->size());
  for (int k = 0; k < rank; ++k)
    SAG_REQUIRE(shape[k] == (*exp)[k], SAG_EINVAL, "sag_load_weight: '%s' dim %d is %lld, expected %lld", tf_name, k, (long long)shape[k], (long long)(*exp)[k]);
  SAG_CHECK_CUDA(cudaSetDevice(h->device));
  DevTensor t;
  t.shape = *exp;
  t.ld = exp->back();
  const int64_t n = numel(*exp);
  SAG_CHECK_CUDA(cudaMalloc(&t.p, sizeof(float) * n));
  cudaError_t e = cudaMemcpy(t.p, host_data, sizeof(float) * n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t.p); set_error("sag_load_weight: copy failed: %s", cudaGetErrorString(e)); return SAG_ECUDA; }
  auto it = h->weights.find(name);
  if (it != h->weights.end()) free_tensor(it->second);
  h->weights[name] = t;
  for (auto it = h->umma.begin(); it != h->umma.end();) {     // stale tensor-core images of this layer
    const std::string scope = name.substr(0, name.rfind('/'));
    if (it->first.compare(0, scope.size() + 1, scope + "#") == 0) { umma_free(&it->second); it = h->umma.erase(it); }
    else ++it;
  }
  if (is_deconv_weight(name)) {
    DevTensor p;
    p.shape = {(*exp)[0] * (*exp)[1], (*exp)[3], (*exp)[2]};
    p.ld = (*exp)[2];
    SAG_CHECK_CUDA(cudaMalloc(&p.p, sizeof(float) * n));
    int r = launch_pack_deconv_weights(t.p, p.p, (int)((*exp)[0] * (*exp)[1]), (int)(*exp)[2], (int)(*exp)[3], 0);
    if (r != SAG_OK) { cudaFree(p.p); return r; }
    SAG_CHECK_CUDA(cudaStreamSynchronize(0));
    auto ip = h->packed.find(name);
    if (ip != h->packed.end()) free_tensor(ip->second);
    h->packed[name] = p;
  }
  h->finalized = 0;
  return SAG_OK;
}

int sag_finalize_weights(sag_handle* h, void* stream) {
  (void)stream;
  SAG_REQUIRE(h != nullptr, SAG_EINVAL, "sag_finalize_weights: NULL handle");
  for (const auto& e : h->expected) {
    // moving statistics are part of the checkpoint layout but are never read by the forward
    // (visual towers run batch-statistics BN: model.py:197, core.py:209-210)
    const std::string& n = e.first;
    bool moving = n.find("/bn/moving_") != std::string::npos;
    SAG_REQUIRE(moving || h->weights.count(n), SAG_ESTATE, "sag_finalize_weights: variable '%s' has not been loaded", n.c_str());
  }
  h->finalized = 1;
  return SAG_OK;
}

// ---- forward --------------------------------------------------------------------------------------------------
size_t sag_workspace_bytes(const sag_handle* h_in, int batch) {
  if (h_in == nullptr || batch <= 0) return 0;
  sag_handle* h = const_cast<sag_handle*>(h_in);
  Arena ar;
  ar.dry = true;
  // once the weights are in place this is also where the batch is PLANNED: the tensor-core operand images whose tile width
  // depends on the row count are packed here (device allocation + a synchronising pack kernel), never inside sag_forward
  ar.prepare = h->finalized != 0 && h->cfg.precision != SAG_PREC_FP32;
  if (ar.prepare && cudaSetDevice(h->device) != cudaSuccess) { set_error("sag_workspace_bytes: cudaSetDevice failed"); return 0; }
  int r = forward(h, nullptr, FrameSrc(), FrameSrc(), nullptr, ar, batch, 0);
  if (r != SAG_OK) return 0;
  return ar.peak + ar.scratch_need + 768;
}

static int forward_checked(sag_handle* h, const float* audio, const FrameSrc& video, const FrameSrc& flow, float* ambix_out, void* workspace,
                           size_t workspace_bytes, int batch, void* stream) {
  SAG_REQUIRE(h != nullptr && audio != nullptr && ambix_out != nullptr && workspace != nullptr, SAG_EINVAL, "sag_forward: NULL argument");
  SAG_REQUIRE(h->finalized, SAG_ESTATE, "sag_forward: call sag_finalize_weights first");
  SAG_REQUIRE(batch > 0, SAG_EINVAL, "sag_forward: batch must be positive");
  Arena dry;
  dry.dry = true;
  SAG_TRY(forward(h, nullptr, FrameSrc(), FrameSrc(), nullptr, dry, batch, 0));
  const size_t need = dry.peak + dry.scratch_need + 768;
  SAG_REQUIRE(need <= workspace_bytes, SAG_ENOMEM, "sag_forward: workspace of %zu bytes is too small, need %zu", workspace_bytes, need);
  Arena ar;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  ar.scratch = reinterpret_cast<float*>(base);                       // split-K region first, then the bump arena
  ar.scratch_cap = dry.scratch_need;
  base = (base + dry.scratch_need + 255) & ~(uintptr_t)255;
  ar.base = reinterpret_cast<char*>(base);
  ar.cap = workspace_bytes - (base - reinterpret_cast<uintptr_t>(workspace));
  return forward(h, audio, video, flow, ambix_out, ar, batch, as_stream(stream));
}

int sag_forward(sag_handle* h, const float* audio, const float* video, const float* flow, float* ambix_out,
                void* workspace, size_t workspace_bytes, int batch, void* stream) {
  return forward_checked(h, audio, FrameSrc(video), FrameSrc(flow), ambix_out, workspace, workspace_bytes, batch, stream);
}

int sag_forward_frames(sag_handle* h, const float* audio, const void* video, int video_format, const void* flow, int flow_format,
                       const double* flow_limits, float* ambix_out, void* workspace, size_t workspace_bytes, int batch, void* stream) {
  SAG_REQUIRE((video_format == SAG_FRAMES_F32 || video_format == SAG_FRAMES_U8) && (flow_format == SAG_FRAMES_F32 || flow_format == SAG_FRAMES_U8),
              SAG_EINVAL, "sag_forward_frames: unknown frame format");
  SAG_REQUIRE(flow == nullptr || flow_format == SAG_FRAMES_F32 || flow_limits != nullptr, SAG_EINVAL,
              "sag_forward_frames: quantised flow frames need their (min, max) limits (feeder.py:147-152)");
  const FrameSrc v(video, video_format == SAG_FRAMES_U8 ? FRAMES_U8_VIDEO : FRAMES_F32, nullptr);
  const FrameSrc f(flow, flow_format == SAG_FRAMES_U8 ? FRAMES_U8_FLOW : FRAMES_F32, flow_limits);
  return forward_checked(h, audio, v, f, ambix_out, workspace, workspace_bytes, batch, stream);
}

int sag_num_tensors(const sag_handle* h) { return h ? (int)h->end_order.size() : SAG_EINVAL; }

int sag_tensor_name(const sag_handle* h, int i, char* buf, int buflen) {
  SAG_REQUIRE(h != nullptr && buf != nullptr && buflen > 0, SAG_EINVAL, "sag_tensor_name: bad argument");
  SAG_REQUIRE(i >= 0 && i < (int)h->end_order.size(), SAG_EINVAL, "sag_tensor_name: index out of range");
  snprintf(buf, buflen, "%s", h->end_order[i].c_str());
  return SAG_OK;
}

int sag_get_tensor(const sag_handle* h, const char* name, const float** dev_ptr, int64_t* shape5, int* rank, int64_t* row_stride) {
  SAG_REQUIRE(h != nullptr && name != nullptr && dev_ptr != nullptr, SAG_EINVAL, "sag_get_tensor: NULL argument");
  auto it = h->ends.find(name);
  SAG_REQUIRE(it != h->ends.end(), SAG_EINVAL, "sag_get_tensor: no tensor named '%s' in the last forward", name);
  *dev_ptr = it->second.p;
  if (rank) *rank = (int)it->second.shape.size();
  if (shape5)
    for (size_t k = 0; k < 5; ++k) shape5[k] = k < it->second.shape.size() ? it->second.shape[k] : 1;
  if (row_stride) *row_stride = it->second.ld;
  return SAG_OK;
}

int sag_get_tensor_format(const sag_handle* h, const char* name, int* format, int64_t* plane_bytes) {
  SAG_REQUIRE(h != nullptr && name != nullptr, SAG_EINVAL, "sag_get_tensor_format: NULL argument");
  auto it = h->ends.find(name);
  SAG_REQUIRE(it != h->ends.end(), SAG_EINVAL, "sag_get_tensor_format: no tensor named '%s' in the last forward", name);
  if (format) *format = it->second.fmt;
  if (plane_bytes) *plane_bytes = it->second.plane;
  return SAG_OK;
}

int sag_last_launch_count(const sag_handle* h) { return h ? h->last_launches : SAG_EINVAL; }

int sag_plan_contraction(int k, int n, int64_t m, int* tile_width, int* k_split) {
  SAG_REQUIRE(k > 0 && n > 0 && m > 0, SAG_EINVAL, "sag_plan_contraction: bad shape %d x %d over %lld rows", k, n, (long long)m);
  if (tile_width) *tile_width = umma_tile_width(k, n, m);
  if (k_split) *k_split = umma_split_k(k, n, m, nullptr);
  return SAG_OK;
}

int sag_plan_stream_k(int k, int n, int64_t m) {
  SAG_REQUIRE(k > 0 && n > 0 && m > 0, SAG_EINVAL, "sag_plan_stream_k: bad shape %d x %d over %lld rows", k, n, (long long)m);
  return umma_stream_k(k, n, m) ? 1 : 0;
}

int sag_stream_k_schedule(int64_t tiles, int k_chunks, int clusters, int cluster, int* items, int max_items) {
  SAG_REQUIRE(tiles > 0 && k_chunks > 0 && clusters > 0 && cluster >= 0 && cluster < clusters && (items != nullptr || max_items == 0),
              SAG_EINVAL, "sag_stream_k_schedule: bad argument");
  return streamk_schedule(tiles, k_chunks, clusters, cluster, items, max_items);
}

int sag_num_profile_records(const sag_handle* h) { return h ? (int)h->prof.recs.size() : SAG_EINVAL; }

int sag_get_profile_record(sag_handle* h, int i, char* name, int name_len, int* category, double* us, double* flops, double* flops_issued,
                           double* bytes, int* tile_width, int* k_split) {
  SAG_REQUIRE(h != nullptr, SAG_EINVAL, "sag_get_profile_record: NULL handle");
  SAG_REQUIRE(i >= 0 && i < (int)h->prof.recs.size(), SAG_EINVAL, "sag_get_profile_record: record %d of %d", i, (int)h->prof.recs.size());
  ProfRec& r = h->prof.recs[i];
  SAG_CHECK_CUDA(cudaEventSynchronize(r.e1));
  float dt = 0.f;
  SAG_CHECK_CUDA(cudaEventElapsedTime(&dt, r.e0, r.e1));
  if (name != nullptr && name_len > 0) snprintf(name, name_len, "%s", r.name);
  if (category) *category = r.cat;
  if (us) *us = dt * 1e3;
  if (flops) *flops = r.flops;
  if (flops_issued) *flops_issued = r.issued;
  if (bytes) *bytes = r.bytes;
  if (tile_width) *tile_width = r.tile;
  if (k_split) *k_split = r.split;
  return SAG_OK;
}

int sag_get_profile(sag_handle* h, int category, double* ms, double* flops, double* bytes, int* launches) {
  SAG_REQUIRE(h != nullptr, SAG_EINVAL, "sag_get_profile: NULL handle");
  SAG_REQUIRE(category >= 0 && category < PROF_NCAT, SAG_EINVAL, "sag_get_profile: category %d outside [0,%d)", category, (int)PROF_NCAT);
  double t = 0, f = 0, b = 0;
  int n = 0;
  static const bool dump = getenv("SAG_PROF_DUMP") != nullptr;   // debug: one line per launch scope of the last forward
  if (dump && category == 0) {
    int i = 0;
    for (auto& r : h->prof.recs) {
      float dt = 0.f;
      cudaEventSynchronize(r.e1);
      cudaEventElapsedTime(&dt, r.e0, r.e1);
      fprintf(stderr, "[prof] %3d cat %d %8.2f us %8.3f GFLOP %8.2f MB\n", i++, r.cat, dt * 1e3, r.flops * 1e-9, r.bytes * 1e-6);
    }
  }
  for (auto& r : h->prof.recs) {
    if (r.cat != category) continue;
    SAG_CHECK_CUDA(cudaEventSynchronize(r.e1));
    float dt = 0.f;
    SAG_CHECK_CUDA(cudaEventElapsedTime(&dt, r.e0, r.e1));
    t += dt; f += r.flops; b += r.bytes; ++n;
  }
  if (ms) *ms = t;
  if (flops) *flops = f;
  if (bytes) *bytes = b;
  if (launches) *launches = n;
  return SAG_OK;
}

// ---- stage entry points --------------------------------------------------------------------------------------
int sag_stft(const float* x, int rows, int n_samples, int wind, int n_overlap, int frame0, int n_frames_out,
             float* cplx_out, int mag0, int n_mag, float* mag_out, void* stream) {
  SAG_REQUIRE(x != nullptr, SAG_EINVAL, "sag_stft: NULL input");
  SAG_REQUIRE(wind > 0 && n_overlap > 0 && wind % n_overlap == 0, SAG_EINVAL, "sag_stft: window %d / overlap %d", wind, n_overlap);
  const int n_winds = n_samples / wind - 1;                              // myutils.py:126
  SAG_REQUIRE(n_winds >= 1, SAG_EINVAL, "sag_stft: %d samples are too few for window %d", n_samples, wind);
  return launch_stft(x, rows, n_samples, wind, wind / n_overlap, n_winds * n_overlap, frame0, n_frames_out, cplx_out,
                     mag0, n_mag, mag_out, as_stream(stream));
}

int sag_istft(const float* cplx_in, int rows, int n_frames, int wind, int n_overlap, float* out, void* stream) {
  SAG_REQUIRE(cplx_in != nullptr && out != nullptr, SAG_EINVAL, "sag_istft: NULL argument");
  SAG_REQUIRE(wind > 0 && n_overlap > 0 && wind % n_overlap == 0, SAG_EINVAL, "sag_istft: window %d / overlap %d", wind, n_overlap);
  const int nf = (n_frames / n_overlap) * n_overlap;
  SAG_REQUIRE(nf > 0, SAG_EINVAL, "sag_istft: needs at least %d frames", n_overlap);
  const int full = (nf / n_overlap) * wind - (n_overlap - 1) * (wind / n_overlap);
  return launch_istft(cplx_in, nullptr, 0, rows, 1, n_frames, wind, n_overlap, 0, full, out, as_stream(stream));
}

int sag_conv2d(const float* x, int n, int h, int w, int cin, const float* w_hwio, int kh, int kw, int cout, int sh,
               int sw, int same_pad, const float* bias, int relu, float* y, int precision, void* stream) {
  SAG_REQUIRE(x != nullptr && w_hwio != nullptr && y != nullptr, SAG_EINVAL, "sag_conv2d: NULL argument");
  GatherGeom g;
  int oh, ow;
  SAG_TRY(make_conv_geom(&g, n, h, w, cin, cin, kh, kw, cout, sh, sw, same_pad, cout, &oh, &ow));
  Epilogue ep{bias, relu, nullptr, nullptr};
  return launch_gather_gemm(precision, x, w_hwio, y, g, ep, as_stream(stream));
}

int sag_deconv2d(const float* x, int n, int h, int w, int cin, const float* w_hwoi, int kh, int kw, int cout, int sh,
                 int sw, const float* bias, int relu, float* y, int precision, void* stream) {
  SAG_REQUIRE(x != nullptr && w_hwoi != nullptr && y != nullptr, SAG_EINVAL, "sag_deconv2d: NULL argument");
  SAG_REQUIRE(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && sh > 0 && sw > 0 && kh > 0 && kw > 0, SAG_EINVAL, "sag_deconv2d: bad dims");
  cudaStream_t st = as_stream(stream);
  if (precision != SAG_PREC_FP32) {               // tcgen05: one sub-pixel GEMM
    const int OH = (h - 1) * sh + kh, OW = (w - 1) * sw + kw;
    GatherGeom g;
    int oh_lim, ow_lim;
    SAG_TRY(make_deconv_subpixel_geom(&g, n, h, w, cin, cin, kh, kw, sh, sw, 0, OH, (int64_t)OH * OW * cout, (int64_t)OW * cout,
                                      cout, 1, &oh_lim, &ow_lim));
    UmmaWeights uw;
    SAG_TRY(umma_pack_deconv(w_hwoi, bias, kh, kw, cout, cin, sh, sw, 0, (int64_t)OW * cout, cout, 1, precision,
                             (int64_t)g.N * g.PH * g.PW, &uw, st));
    int r = SAG_OK;
    g.Cout = uw.N;
    Epilogue ep{bias, relu, nullptr, nullptr};
    if (r == SAG_OK) r = launch_gather_gemm_umma(x, uw, y, g, ep, oh_lim, ow_lim, nullptr, st);
    cudaStreamSynchronize(st);
    umma_free(&uw);
    return r;
  }
  float* wp = nullptr;
  SAG_CHECK_CUDA(cudaMallocAsync(&wp, sizeof(float) * (size_t)kh * kw * cin * cout, st));
  int r = launch_pack_deconv_weights(w_hwoi, wp, kh * kw, cout, cin, st);
  const int OH = (h - 1) * sh + kh, OW = (w - 1) * sw + kw;
  Epilogue ep{bias, relu, nullptr, nullptr};
  for (int py = 0; py < sh && r == SAG_OK; ++py)
    for (int px = 0; px < sw && r == SAG_OK; ++px) {
      GatherGeom g;
      int q = make_deconv_phase_geom(&g, n, h, w, cin, cin, kh, kw, cout, sh, sw, py, px, 0, OH, (int64_t)OH * OW * cout,
                                     (int64_t)OW * cout, cout, 1);
      if (q == 1) continue;
      r = q;
      if (r == SAG_OK) r = launch_gather_gemm(precision, x, wp, y, g, ep, st);
    }
  cudaFreeAsync(wp, st);
  return r;
}

int sag_fc(const float* x, int rows, int in, const float* w, int out, const float* bias, int relu, float* y,
           int precision, void* stream) {
  SAG_REQUIRE(x != nullptr && w != nullptr && y != nullptr, SAG_EINVAL, "sag_fc: NULL argument");
  GatherGeom g;
  int oh, ow;
  SAG_TRY(make_conv_geom(&g, 1, 1, rows, in, in, 1, 1, out, 1, 1, 0, out, &oh, &ow));
  Epilogue ep{bias, relu, nullptr, nullptr};
  return launch_gather_gemm(precision, x, w, y, g, ep, as_stream(stream));
}

int sag_batchnorm_train(const float* x, int64_t rows, int c, const float* gamma, const float* beta,
                        const float* residual, int relu, float* y, void* scratch, void* stream) {
  SAG_REQUIRE(x != nullptr && gamma != nullptr && beta != nullptr && y != nullptr && scratch != nullptr, SAG_EINVAL, "sag_batchnorm_train: NULL argument");
  SAG_REQUIRE(rows > 0 && c > 0, SAG_EINVAL, "sag_batchnorm_train: bad dims");
  cudaStream_t st = as_stream(stream);
  unsigned long long* sum = reinterpret_cast<unsigned long long*>(scratch);      // [c][2] + [c][2] fixed-point words = 4*c*8 bytes
  unsigned long long* sqs = sum + 2 * c;
  SAG_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(unsigned long long) * 4 * c, st));
  SAG_TRY(launch_channel_stats(x, rows, c, sum, sqs, st));
  BnStats bn;
  bn.sum = sum; bn.sqs = sqs; bn.gamma = gamma; bn.beta = beta;
  bn.inv_count = 1.0 / (double)rows;
  bn.eps = 1e-3f;
  return launch_bn_apply_stats(x, bn, ActView(residual), relu, ActView(y), rows, c, st);
}

int sag_maxpool_3x3s2_same(const float* x, int n, int h, int w, int c, float* y, void* stream) {
  SAG_REQUIRE(x != nullptr && y != nullptr, SAG_EINVAL, "sag_maxpool: NULL argument");
  return launch_bn_relu_maxpool(x, nullptr, nullptr, n, h, w, c, y, as_stream(stream));
}

int sag_resnet18(sag_handle* h, const char* scope, const float* x, int batch, float* y, void* workspace,
                 size_t workspace_bytes, void* stream) {
  SAG_REQUIRE(h != nullptr && scope != nullptr && x != nullptr && y != nullptr && workspace != nullptr, SAG_EINVAL, "sag_resnet18: NULL argument");
  SAG_REQUIRE(batch > 0, SAG_EINVAL, "sag_resnet18: batch must be positive");
  Arena dry;
  dry.dry = true;
  dry.prepare = h->cfg.precision != SAG_PREC_FP32;      // stage entry point: builds the operand images it needs itself
  Act yact;
  SAG_TRY(resnet18_tower(h, scope, FrameSrc(), batch, h->cfg.frame_h, h->cfg.frame_w, &yact, dry, 0));
  const size_t need = dry.peak + dry.scratch_need + 768;
  SAG_REQUIRE(need <= workspace_bytes, SAG_ENOMEM, "sag_resnet18: workspace of %zu bytes is too small, need %zu", workspace_bytes, need);
  Arena ar;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  ar.scratch = reinterpret_cast<float*>(base);
  ar.scratch_cap = dry.scratch_need;
  base = (base + dry.scratch_need + 255) & ~(uintptr_t)255;
  ar.base = reinterpret_cast<char*>(base);
  ar.cap = workspace_bytes - (base - reinterpret_cast<uintptr_t>(workspace));
  h->ends.clear();
  h->end_order.clear();
  SAG_TRY(resnet18_tower(h, scope, FrameSrc(x), batch, h->cfg.frame_h, h->cfg.frame_w, &yact, ar, as_stream(stream)));
  const int fh = (h->cfg.frame_h + 31) / 32, fw = (h->cfg.frame_w + 31) / 32;
  return launch_act_to_f32(yact.v, y, (int64_t)batch * fh * fw * 512, as_stream(stream));
}

int sag_mix(const float* x_sep, const float* loc, int batch, int tracks, int t, int segments, float* out, void* stream) {
  SAG_REQUIRE(x_sep != nullptr && loc != nullptr && out != nullptr, SAG_EINVAL, "sag_mix: NULL argument");
  SAG_REQUIRE(batch > 0 && tracks > 0 && t > 0, SAG_EINVAL, "sag_mix: bad dims");
  return launch_mix(x_sep, loc, batch, tracks, t, segments, out, as_stream(stream));
}

size_t sag_metrics_scratch_bytes(int batch, int t) { (void)batch; (void)t; return 256; }

int sag_metrics(const float* pred, const float* gt, int batch, int t, int audio_rate, float* stft_ps, float* lsd_ps,
                float* mse_ps, float* snr_ps, float* env_ps, float* amp, void* scratch, void* stream) {
  SAG_REQUIRE(pred != nullptr && gt != nullptr && stft_ps != nullptr && lsd_ps != nullptr && mse_ps != nullptr &&
              snr_ps != nullptr && amp != nullptr, SAG_EINVAL, "sag_metrics: NULL argument");
  return launch_metrics(pred, gt, batch, t, audio_rate, stft_ps, lsd_ps, mse_ps, snr_ps, env_ps, amp, scratch, as_stream(stream));
}

int sag_mel_lsd(const float* pred, const float* gt, int batch, int t, int audio_rate, float* mel_lsd_ps, void* stream) {
  SAG_REQUIRE(pred != nullptr && gt != nullptr && mel_lsd_ps != nullptr, SAG_EINVAL, "sag_mel_lsd: NULL argument");
  return launch_mel_lsd(pred, gt, batch, t, audio_rate, mel_lsd_ps, as_stream(stream));
}

int sag_sh_rms_dims(float ang_res, int* n_nu, int* n_phi) {
  SAG_REQUIRE(n_nu != nullptr && n_phi != nullptr, SAG_EINVAL, "sag_sh_rms_dims: NULL argument");
  return sh_mesh_dims(ang_res, n_nu, n_phi);
}

int sag_sh_rms(const float* ambi, int batch, int t, float ang_res, float* rms, void* stream) {
  SAG_REQUIRE(ambi != nullptr && rms != nullptr, SAG_EINVAL, "sag_sh_rms: NULL argument");
  return launch_sh_rms(ambi, batch, t, ang_res, rms, as_stream(stream));
}

}  // extern "C"
