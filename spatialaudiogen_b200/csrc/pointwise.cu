// HBM-bound pointwise / reduction kernels: batch-statistics BN (reference core.py:209-210),
// 3x3/2 SAME max-pool (resnet.py:135), tf.tile replacements (model.py:232,262,295) and the time-varying
// 32->3 mixing of model.py:424-432.  All are streaming kernels: float4 accesses, grid sized in multiples
// of the SM count, one pass over the data.
#include "common.cuh"
#include <cuda_bf16.h>

namespace sag {

// ---- split-bf16 activation stores / loads (ACT_BF2: hi = bf16(x), lo = bf16(x - hi)) ----
__device__ __forceinline__ void store_act4(void* p, int fmt, int64_t plane, int64_t e, const float4& v) {
  if (fmt == ACT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + e) = v;
    return;
  }
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + e) =
      make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  if (plane != 0) {
    __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
    __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(p) + plane) + e) =
        make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
  }
}
__device__ __forceinline__ float4 load_act4(const void* p, int fmt, int64_t plane, int64_t e) {
  if (fmt == ACT_F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + e));
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + e));
  float4 v = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xffff0000u), __uint_as_float(h.y << 16),
                         __uint_as_float(h.y & 0xffff0000u));
  if (plane != 0) {
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(
        reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const char*>(p) + plane) + e));
    v.x += __uint_as_float(l.x << 16); v.y += __uint_as_float(l.x & 0xffff0000u);
    v.z += __uint_as_float(l.y << 16); v.w += __uint_as_float(l.y & 0xffff0000u);
  }
  return v;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// ---- batch-norm consumers: scale / shift are derived from the raw batch statistics in the block prologue ----
//      scale = gamma*rsqrt(var+eps), shift = beta - mean*scale (biased variance); y = act(x*scale + shift [+ res])
__device__ __forceinline__ void bn_scale_shift_to_smem(const BnStats& bn, int c, float* s_scale, float* s_shift) {
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    const double mean = fx_value(bn.sum + 2 * i) * bn.inv_count;
    double var = fx_value(bn.sqs + 2 * i) * bn.inv_count - mean * mean;
    if (var < 0) var = 0;
    const double sc = (double)__ldg(bn.gamma + i) / sqrt(var + (double)bn.eps);
    s_scale[i] = (float)sc;
    s_shift[i] = (float)((double)__ldg(bn.beta + i) - mean * sc);
  }
  __syncthreads();
}

// Memory-level parallelism: a thread fetches BN_UNROLL vectors (and their residuals) before it touches any of them -- with one
// 16-byte load in flight per thread the kernel sat at 35-44 % of the HBM rate on scoreboard stalls (profiles/r1_small_kernels_ncu.txt).
constexpr int BN_UNROLL = 4;
template <bool BN_STREAM>
__global__ void __launch_bounds__(256, 4) bn_apply_stats_kernel(const float4* __restrict__ x, const BnStats bn, const ActView res, int relu,
                                                                const ActView y, int64_t n4, int c) {
  extern __shared__ float s_ss[];
  float* s_scale = s_ss;
  float* s_shift = s_ss + c;
  pdl_prologue();
  bn_scale_shift_to_smem(bn, c, s_scale, s_shift);
  const int c4 = c / 4;
  const bool has_res = res.p != nullptr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += BN_UNROLL * stride) {
    float4 v[BN_UNROLL], r[BN_UNROLL];
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {                    // vector i0 + u*stride: all loads first
      const int64_t i = i0 + (int64_t)u * stride;
      if (i < n4) {
        v[u] = BN_STREAM ? __ldcs(x + i) : __ldg(x + i);      // raw conv output: dead after this read (evict-first)
        if (has_res) r[u] = load_act4(res.p, res.fmt, res.plane, i * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) {
      const int64_t i = i0 + (int64_t)u * stride;
      if (i < n4) {
        const int cc = (int)(i % c4) * 4;
        const float4 sc = *reinterpret_cast<const float4*>(s_scale + cc);
        const float4 sh = *reinterpret_cast<const float4*>(s_shift + cc);
        float4 w = v[u];
        w.x = fmaf(w.x, sc.x, sh.x); w.y = fmaf(w.y, sc.y, sh.y); w.z = fmaf(w.z, sc.z, sh.z); w.w = fmaf(w.w, sc.w, sh.w);
        if (has_res) { w.x += r[u].x; w.y += r[u].y; w.z += r[u].z; w.w += r[u].w; }
        if (relu) { w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f); }
        store_act4(y.p, y.fmt, y.plane, i * 4, w);
      }
    }
  }
}

int launch_bn_apply_stats(const float* x, const BnStats& bn, const ActView& residual, int relu, const ActView& y, int64_t rows,
                          int c, cudaStream_t st) {
  SAG_REQUIRE(c % 4 == 0 && c <= 4096, SAG_EINVAL, "bn_apply: channels %d not a multiple of 4 (or too many)", c);
  int64_t n4 = rows * c / 4;
  if (n4 == 0) return SAG_OK;
  // one co-resident wave (4 blocks of 256 threads per SM, launch bounds): every block pays the scale / shift prologue once
  // and streams BN_UNROLL vectors per thread per round; small tensors get exactly one round per thread
  int64_t blocks = cdiv64(n4, 256 * BN_UNROLL);
  int64_t cap = (int64_t)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  static const bool stream = [] { const char* v = getenv("SAG_BN_STREAM"); return v == nullptr || atoi(v) != 0; }();
  if (stream)
    launch_pdl(bn_apply_stats_kernel<true>, dim3((unsigned)blocks), dim3(256), 2 * c * sizeof(float), st, reinterpret_cast<const float4*>(x), bn,
               residual, relu, y, n4, c);
  else
    launch_pdl(bn_apply_stats_kernel<false>, dim3((unsigned)blocks), dim3(256), 2 * c * sizeof(float), st, reinterpret_cast<const float4*>(x), bn,
               residual, relu, y, n4, c);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- fused BN + ReLU + max-pool 3x3/2 SAME (pad before 0: TF puts the odd pad cell after) ----
__global__ void bn_relu_maxpool_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                       const float* __restrict__ shift, int n, int h, int w, int c, int oh, int ow,
                                       int pt, int pl, const ActView y) {
  int c4 = c / 4;
  int64_t total = (int64_t)n * oh * ow * c4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int cc = (int)(i % c4) * 4;
    int64_t r = i / c4;
    int ox = (int)(r % ow); r /= ow;
    int oy = (int)(r % oh);
    int b = (int)(r / oh);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale != nullptr) {
      sc = __ldg(reinterpret_cast<const float4*>(scale + cc));
      sh = __ldg(reinterpret_cast<const float4*>(shift + cc));
    }
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      int iy = oy * 2 + dy - pt;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        int ix = ox * 2 + dx - pl;
        if (ix < 0 || ix >= w) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * h + iy) * w + ix) * c + cc));
        if (scale != nullptr) {
          v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
          v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
        }
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    store_act4(y.p, y.fmt, y.plane, i * 4, m);
  }
}

// The nine taps are fetched branch-free: an out-of-range tap is clamped onto the nearest row / column, which is itself a tap of the
// same window, so the maximum is unchanged (max is idempotent) and the nine 16-byte loads are all in flight before the first use --
// the skip-by-branch form issued them one basic block at a time and ran at a third of the HBM rate.
__global__ void __launch_bounds__(256, 4) bn_relu_maxpool_stats_kernel(const float* __restrict__ x, const BnStats bn, int n, int h, int w,
                                                                       int c, int oh, int ow, int pt, int pl, const ActView y) {
  extern __shared__ float s_ss[];
  float* s_scale = s_ss;
  float* s_shift = s_ss + c;
  pdl_prologue();
  bn_scale_shift_to_smem(bn, c, s_scale, s_shift);
  int c4 = c / 4;
  int64_t total = (int64_t)n * oh * ow * c4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int cc = (int)(i % c4) * 4;
    int64_t r = i / c4;
    int ox = (int)(r % ow); r /= ow;
    int oy = (int)(r % oh);
    int b = (int)(r / oh);
    float4 v[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = min(max(oy * 2 + dy - pt, 0), h - 1);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = min(max(ox * 2 + dx - pl, 0), w - 1);
        v[dy * 3 + dx] = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * h + iy) * w + ix) * c + cc));
      }
    }
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + cc);
    const float4 sh = *reinterpret_cast<const float4*>(s_shift + cc);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      m.x = fmaxf(m.x, fmaxf(fmaf(v[t].x, sc.x, sh.x), 0.f)); m.y = fmaxf(m.y, fmaxf(fmaf(v[t].y, sc.y, sh.y), 0.f));
      m.z = fmaxf(m.z, fmaxf(fmaf(v[t].z, sc.z, sh.z), 0.f)); m.w = fmaxf(m.w, fmaxf(fmaf(v[t].w, sc.w, sh.w), 0.f));
    }
    store_act4(y.p, y.fmt, y.plane, i * 4, m);
  }
}

int launch_bn_relu_maxpool_stats(const float* x, const BnStats& bn, int n, int h, int w, int c, const ActView& y, cudaStream_t st) {
  SAG_REQUIRE(c % 4 == 0 && c <= 4096, SAG_EINVAL, "maxpool: channels %d not a multiple of 4 (or too many)", c);
  int oh, ow;
  int pt = same_pad_before(h, 3, 2, &oh), pl = same_pad_before(w, 3, 2, &ow);
  int64_t total = (int64_t)n * oh * ow * (c / 4);
  int64_t blocks = cdiv64(total, 256 * 2);
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_pdl(bn_relu_maxpool_stats_kernel, dim3((unsigned)blocks), dim3(256), 2 * c * sizeof(float), st, x, bn, n, h, w, c, oh, ow, pt, pl, y);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

int launch_bn_relu_maxpool(const float* x, const float* scale, const float* shift, int n, int h, int w, int c,
                           const ActView& y, cudaStream_t st) {
  SAG_REQUIRE(c % 4 == 0, SAG_EINVAL, "maxpool: channels %d not a multiple of 4", c);
  int oh, ow;
  int pt = same_pad_before(h, 3, 2, &oh), pl = same_pad_before(w, 3, 2, &ow);
  int64_t total = (int64_t)n * oh * ow * (c / 4);
  int64_t blocks = cdiv64(total, 256);
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  bn_relu_maxpool_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, scale, shift, n, h, w, c, oh, ow, pt, pl, y);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- per-channel sum / sum of squares over rows (stand-alone BN statistics) ----
// block = 256 threads = 8 row lanes x 32 channel lanes(x4 via float4 when possible): generic scalar version
__global__ void channel_stats_kernel(const float* __restrict__ x, int64_t rows, int c, unsigned long long* __restrict__ sum,
                                     unsigned long long* __restrict__ sqs) {
  // each block handles a contiguous slab of rows; thread t covers channels t, t+blockDim, ...
  int64_t rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      float v = __ldg(x + r * c + ch);
      s += v;
      q = fmaf(v, v, q);
    }
    if (r1 > r0) {
      fx_atomic_add(sum + 2 * ch, s);
      fx_atomic_add(sqs + 2 * ch, q);
    }
  }
}

int launch_channel_stats(const float* x, int64_t rows, int c, unsigned long long* sum, unsigned long long* sqs, cudaStream_t st) {
  int threads = c >= 256 ? 256 : (c >= 128 ? 128 : (c >= 64 ? 64 : 32));
  int64_t blocks = cdiv64(rows, 64);
  int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  channel_stats_kernel<<<(unsigned)blocks, threads, 0, st>>>(x, rows, c, sum, sqs);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- split-bf16 (or fp32) activation -> dense fp32 (stage entry points / taps); n % 4 == 0 ----
__global__ void act_to_f32_kernel(const ActView src, float4* __restrict__ dst, int64_t n4) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    dst[i] = load_act4(src.p, src.fmt, src.plane, i * 4);
}

int launch_act_to_f32(const ActView& src, float* dst, int64_t n, cudaStream_t st) {
  SAG_REQUIRE(n % 4 == 0, SAG_EINVAL, "act_to_f32: %lld elements not a multiple of 4", (long long)n);
  int64_t blocks = cdiv64(n / 4, 256);
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return SAG_OK;
  act_to_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, reinterpret_cast<float4*>(dst), n / 4);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- tf.tile replacement: dst[(g*reps + r)*dst_ld + 0..c) = src[g*src_ld + 0..c) ----
template <class T>
__global__ void tile_rows_kernel(const T* __restrict__ src, int64_t src_ld, T* __restrict__ dst, int64_t dst_ld, int groups,
                                 int reps, int c) {
  pdl_prologue();
  int64_t total = (int64_t)groups * reps * c;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int ch = (int)(i % c);
    int64_t row = i / c;
    int64_t gidx = row / reps;
    dst[row * dst_ld + ch] = src[gidx * src_ld + ch];
  }
}

int launch_tile_rows(const ActView& src, int64_t src_ld, const ActView& dst, int64_t dst_ld, int groups, int reps, int c,
                     cudaStream_t st) {
  int64_t total = (int64_t)groups * reps * c;
  if (total == 0) return SAG_OK;
  SAG_REQUIRE(src.fmt == dst.fmt && (src.plane != 0) == (dst.plane != 0), SAG_EINVAL, "tile_rows: format mismatch");
  int64_t blocks = cdiv64(total, 256);
  int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (src.fmt == ACT_F32) {
    launch_pdl(tile_rows_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, st, reinterpret_cast<const float*>(src.p), src_ld,
               reinterpret_cast<float*>(dst.p), dst_ld, groups, reps, c);
    SAG_LAUNCH_CHECK();
  } else {
    for (int pl = 0; pl < (src.plane != 0 ? 2 : 1); ++pl) {
      launch_pdl(tile_rows_kernel<unsigned short>, dim3((unsigned)blocks), dim3(256), 0, st,
                 reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(src.p) + pl * src.plane), src_ld,
                 reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(dst.p) + pl * dst.plane), dst_ld, groups, reps, c);
      SAG_LAUNCH_CHECK();
    }
  }
  return SAG_OK;
}

// ---- decode (model.py:424-432): out[b,n,o] = sum_k loc[b,seg(n),o*(K+1)+k]*x_sep[b,k,n] + loc[b,seg(n),o*(K+1)+K]
// loc weights are piecewise constant over t/segments samples (model.py:262-263 tiles them); they stay in smem.
// One block per (b, chunk of 256 samples); coalesced reads along n for every track.
__global__ void mix_kernel(const float* __restrict__ x_sep, const float* __restrict__ loc, int tracks, int t,
                           int segments, float* __restrict__ out) {
  extern __shared__ float s_loc[];   // [3*(tracks+1)]
  pdl_prologue();
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * blockDim.x;
  const int seg_len = t / segments;
  const int K1 = tracks + 1;
  const int n = n0 + threadIdx.x;
  // a block never straddles a segment boundary when seg_len % blockDim.x == 0; otherwise reload per thread
  const bool uniform = (seg_len % blockDim.x) == 0;
  int seg = min(n0 / seg_len, segments - 1);
  if (uniform) {
    for (int i = threadIdx.x; i < 3 * K1; i += blockDim.x) s_loc[i] = __ldg(loc + ((int64_t)b * segments + seg) * 3 * K1 + i);
    __syncthreads();
  }
  if (n >= t) return;
  const float* lw = s_loc;
  if (!uniform) {
    seg = min(n / seg_len, segments - 1);
    lw = loc + ((int64_t)b * segments + seg) * 3 * K1;
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  const float* xs = x_sep + (int64_t)b * tracks * t + n;
  for (int k = 0; k < tracks; ++k) {
    float v = __ldg(xs + (int64_t)k * t);
    a0 = fmaf(lw[k], v, a0);
    a1 = fmaf(lw[K1 + k], v, a1);
    a2 = fmaf(lw[2 * K1 + k], v, a2);
  }
  float* o = out + ((int64_t)b * t + n) * 3;
  o[0] = a0 + lw[tracks];
  o[1] = a1 + lw[K1 + tracks];
  o[2] = a2 + lw[2 * K1 + tracks];
}

int launch_mix(const float* x_sep, const float* loc, int batch, int tracks, int t, int segments, float* out,
               cudaStream_t st) {
  SAG_REQUIRE(segments > 0 && t % segments == 0, SAG_EINVAL, "mix: %d samples not divisible into %d segments", t, segments);
  int threads = 64;
  dim3 grid(cdiv(t, threads), batch);
  launch_pdl(mix_kernel, grid, dim3(threads), 3 * (tracks + 1) * sizeof(float), st, x_sep, loc, tracks, t, segments, out);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- frame ingest: (n,h,w,c) -> 2x2 space-to-depth of the zero-bordered image, 16 channels per pixel ----
// The source is either the prepared fp32 frame or the uint8 frame as decoded from the jpg, prepared here:
//   FRAMES_U8 video: myutils.img_prep_fcn (myutils.py:88-89)  x/255. - 0.5, evaluated in double like numpy and rounded once to
//                    fp32 (what feeding the float64 array into the float32 placeholder does) -- a 256-entry table per block;
//   FRAMES_U8 flow : FlowReader.get_by_index (feeder.py:147-161): channel 2 = magnitude de-quantised with the frame's
//                    (min, max) limits (float32 chunk x float64 limits, rounded to fp32 after each in-place step), channel 0 =
//                    angle * 2pi/255 in fp32, result (mag cos, mag sin, mag).
template <int C, int KIND>
__device__ __forceinline__ void load_frame_pixel(const void* __restrict__ x, int64_t pix, const float* lut, double lim_scale, double lim_min,
                                                 float* v) {
  if (KIND == FRAMES_F32) {
    const float* p = reinterpret_cast<const float*>(x) + pix * C;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) v[ch] = __ldg(p + ch);
  } else {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(x) + pix * C;
    unsigned char u[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) u[ch] = __ldg(p + ch);
    if (KIND == FRAMES_U8_VIDEO) {
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = lut[u[ch]];
    } else if (KIND == FRAMES_U8_VIDEO_INT) {
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = (float)(2 * (int)u[ch] - 255);      // odd integers up to 255: exact in bf16
    } else {
      float m = (float)((double)(float)u[C - 1] * lim_scale);       // chunk[..., 2] *= (m_max - m_min) / 255.
      m = (float)((double)m + lim_min);                              // chunk[..., 2] += m_min
      const float ang = (float)u[0] * (float)(2.0 * 3.14159265358979323846 / 255.0);
      float sn, cs;
      sincosf(ang, &sn, &cs);
      v[0] = m * cs;
      if (C > 1) v[1] = m * sn;
      if (C > 2) v[C - 1] = m;
    }
  }
}
__device__ __forceinline__ void frame_lut_to_smem(float* lut) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = (float)((double)i / 255.0 - 0.5);
  __syncthreads();
}

template <int C, int KIND>
__global__ void space_to_depth16_kernel(const void* __restrict__ x, const double* __restrict__ lims, int n, int h, int w, int pt, int pl,
                                        int h2, int w2, const ActView out) {
  __shared__ float s_lut[256];
  pdl_prologue();
  if (KIND == FRAMES_U8_VIDEO) frame_lut_to_smem(s_lut);
  const int64_t total = (int64_t)n * h2 * w2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int x2 = (int)(i % w2);
    const int64_t r = i / w2;
    const int y2 = (int)(r % h2);
    const int b = (int)(r / h2);
    double lscale = 0.0, lmin = 0.0;
    if (KIND == FRAMES_U8_FLOW) { lmin = lims[2 * b]; lscale = (lims[2 * b + 1] - lmin) / 255.0; }
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      // border pixels read a clamped (valid) address and are zeroed by a select: no branch between the 4*C loads
      const int iy = 2 * y2 + (sub >> 1) - pt, ix = 2 * x2 + (sub & 1) - pl;
      const bool inside = (unsigned)iy < (unsigned)h && (unsigned)ix < (unsigned)w;
      float t[C];
      load_frame_pixel<C, KIND>(x, ((int64_t)b * h + min(max(iy, 0), h - 1)) * w + min(max(ix, 0), w - 1), s_lut, lscale, lmin, t);
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[sub * C + ch] = inside ? t[ch] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 16; e += 4) store_act4(out.p, out.fmt, out.plane, i * 16 + e, make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
  }
}

template <int C>
static void launch_s2d_kind(int kind, unsigned blocks, cudaStream_t st, const void* x, const double* lims, int n, int h, int w, int pt, int pl,
                            int h2, int w2, const ActView& out) {
  if (kind == FRAMES_U8_VIDEO) launch_pdl(space_to_depth16_kernel<C, FRAMES_U8_VIDEO>, dim3(blocks), dim3(256), 0, st, x, lims, n, h, w, pt, pl, h2, w2, out);
  else if (kind == FRAMES_U8_VIDEO_INT) launch_pdl(space_to_depth16_kernel<C, FRAMES_U8_VIDEO_INT>, dim3(blocks), dim3(256), 0, st, x, lims, n, h, w, pt, pl, h2, w2, out);
  else if (kind == FRAMES_U8_FLOW) launch_pdl(space_to_depth16_kernel<C, FRAMES_U8_FLOW>, dim3(blocks), dim3(256), 0, st, x, lims, n, h, w, pt, pl, h2, w2, out);
  else launch_pdl(space_to_depth16_kernel<C, FRAMES_F32>, dim3(blocks), dim3(256), 0, st, x, lims, n, h, w, pt, pl, h2, w2, out);
}

int launch_space_to_depth16(const FrameSrc& src, int n, int h, int w, int c, int pt, int pl, int h2, int w2, const ActView& out,
                            cudaStream_t st) {
  SAG_REQUIRE(c >= 1 && 4 * c <= 16, SAG_EINVAL, "space_to_depth16: %d channels", c);
  SAG_REQUIRE(src.kind != FRAMES_U8_FLOW || (c == 3 && src.lims != nullptr), SAG_EINVAL, "quantised flow frames need 3 channels and their limits");
  const int64_t total = (int64_t)n * h2 * w2;
  int64_t blocks = cdiv64(total, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  switch (c) {
    case 1: launch_s2d_kind<1>(src.kind == FRAMES_U8_FLOW ? FRAMES_F32 : src.kind, (unsigned)blocks, st, src.p, src.lims, n, h, w, pt, pl, h2, w2, out); break;
    case 2: launch_s2d_kind<2>(src.kind == FRAMES_U8_FLOW ? FRAMES_F32 : src.kind, (unsigned)blocks, st, src.p, src.lims, n, h, w, pt, pl, h2, w2, out); break;
    case 3: launch_s2d_kind<3>(src.kind, (unsigned)blocks, st, src.p, src.lims, n, h, w, pt, pl, h2, w2, out); break;
    default: launch_s2d_kind<4>(src.kind == FRAMES_U8_FLOW ? FRAMES_F32 : src.kind, (unsigned)blocks, st, src.p, src.lims, n, h, w, pt, pl, h2, w2, out); break;
  }
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// uint8 frames -> the prepared fp32 frames (exact-fp32 FFMA path, stage entry points): same arithmetic as above
template <int KIND>
__global__ void frames_to_f32_kernel(const void* __restrict__ x, const double* __restrict__ lims, int64_t pixels_per_image, int64_t total,
                                     float* __restrict__ out) {
  __shared__ float s_lut[256];
  pdl_prologue();
  if (KIND == FRAMES_U8_VIDEO) frame_lut_to_smem(s_lut);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    double lscale = 0.0, lmin = 0.0;
    if (KIND == FRAMES_U8_FLOW) { const int64_t b = i / pixels_per_image; lmin = lims[2 * b]; lscale = (lims[2 * b + 1] - lmin) / 255.0; }
    float t[3];
    load_frame_pixel<3, KIND>(x, i, s_lut, lscale, lmin, t);
    out[i * 3] = t[0]; out[i * 3 + 1] = t[1]; out[i * 3 + 2] = t[2];
  }
}

int launch_frames_to_f32(const FrameSrc& src, int n, int h, int w, float* out, cudaStream_t st) {
  SAG_REQUIRE(src.kind == FRAMES_U8_VIDEO || (src.kind == FRAMES_U8_FLOW && src.lims != nullptr), SAG_EINVAL, "frames_to_f32: bad source");
  const int64_t total = (int64_t)n * h * w;
  int64_t blocks = cdiv64(total, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (src.kind == FRAMES_U8_VIDEO) launch_pdl(frames_to_f32_kernel<FRAMES_U8_VIDEO>, dim3((unsigned)blocks), dim3(256), 0, st, src.p, src.lims, (int64_t)h * w, total, out);
  else launch_pdl(frames_to_f32_kernel<FRAMES_U8_FLOW>, dim3((unsigned)blocks), dim3(256), 0, st, src.p, src.lims, (int64_t)h * w, total, out);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}


// ---- tf.nn.conv2d_transpose weights [kh,kw,Cout,Cin] (core.py:118) -> per-tap [Cin][Cout] slabs ----
__global__ void pack_deconv_w_kernel(const float* __restrict__ w, float* __restrict__ out, int taps, int cout, int cin) {
  int64_t total = (int64_t)taps * cout * cin;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int co = (int)(i % cout);
    int64_t r = i / cout;
    int ci = (int)(r % cin);
    int t = (int)(r / cin);
    out[i] = __ldg(w + ((int64_t)t * cout + co) * cin + ci);
  }
}

int launch_pack_deconv_weights(const float* w_hwoi, float* out, int taps, int cout, int cin, cudaStream_t st) {
  int64_t total = (int64_t)taps * cout * cin;
  if (total == 0) return SAG_OK;
  int64_t blocks = cdiv64(total, 256);
  int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  pack_deconv_w_kernel<<<(unsigned)blocks, 256, 0, st>>>(w_hwoi, out, taps, cout, cin);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

}  // namespace sag
