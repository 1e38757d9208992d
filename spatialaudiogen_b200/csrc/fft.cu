// Framed STFT / masked inverse STFT with overlap-add, as fused shared-memory Stockham FFT kernels.
//
//   launch_stft : myutils.stft (reference myutils.py:119-147)  frame t = samples [hop*t, hop*t+wind) x periodic Hann,
//                 unnormalised two-sided forward FFT (tf.fft), plus tf.abs (model.py:178) for the encoder frames.
//   launch_istft: sigmoid mask x STFT (model.py:334-337) -> real(tf.ifft) -> de-interleave/trim/sum/n_overlap
//                 (myutils.py:181-211) -> crop (model.py:344-347), one CTA per (window, track).
//
// Both are HBM-bound: the whole transform lives in shared memory, the audio stream / mask is read once with
// coalesced accesses, results are written once.  Twiddles and the Hann window are fp32 tables rounded from
// float64 exactly as the reference builds them (np.cos in float64 -> tf.constant(float32)).
#include "common.cuh"
#include "fft_device.cuh"
#include <cuda_bf16.h>
#include <mutex>
#include <cmath>

namespace sag {

static std::mutex g_plan_mu;
static std::map<std::pair<int, int>, FftPlan> g_plans;   // (device, n) -> plan

int get_plan(int n, FftPlan* out) {
  int dev = 0;
  SAG_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto it = g_plans.find({dev, n});
  if (it != g_plans.end()) { *out = it->second; return SAG_OK; }
  FftPlan p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  int m = n;
  auto add_pass = [&](unsigned r) { p.radix_code |= r << (4 * p.npass); ++p.npass; };
  while (m % 4 == 0 && p.npass < kMaxPasses) { add_pass(4); m /= 4; }
  while (m % 2 == 0 && p.npass < kMaxPasses) { add_pass(2); m /= 2; }
  while (m % 3 == 0 && p.npass < kMaxPasses) { add_pass(3); m /= 3; }
  while (m % 5 == 0 && p.npass < kMaxPasses) { add_pass(5); m /= 5; }
  SAG_REQUIRE(m == 1, SAG_EUNSUPPORTED, "fft: length %d is not a product of 2,3,5 (or too many passes)", n);
  SAG_REQUIRE(n <= 8192, SAG_EUNSUPPORTED, "fft: length %d too large for the shared-memory transform", n);
  std::vector<float2> tw(n);
  std::vector<float> hw(n);
  for (int i = 0; i < n; ++i) {
    double a = -2.0 * M_PI * (double)i / (double)n;
    tw[i] = make_float2((float)cos(a), (float)sin(a));
    hw[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI / (double)n * (double)i));   // myutils.py:134
  }
  float2* dtw = nullptr;
  float* dh = nullptr;
  SAG_CHECK_CUDA(cudaMalloc(&dtw, sizeof(float2) * n));
  SAG_CHECK_CUDA(cudaMalloc(&dh, sizeof(float) * n));
  SAG_CHECK_CUDA(cudaMemcpy(dtw, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
  SAG_CHECK_CUDA(cudaMemcpy(dh, hw.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
  p.tw = dtw;
  p.hann = dh;
  g_plans[{dev, n}] = p;
  *out = p;
  return SAG_OK;
}

int fft_prepare(int n) {   // create tables ahead of time (outside stream capture)
  FftPlan p;
  return get_plan(n, &p);
}

// ---- K1: frame + window + FFT (+ magnitude) --------------------------------------------------------------------
// grid (ceil(n_frames_launch / 2), rows): two consecutive frames per CTA, transformed side by side (one set of block-wide
// passes and twiddle fetches).  They are NOT packed into one complex transform: an all-zero frame must come out exactly
// zero (frame indexing is checked bit-exactly), which the Hermitian split of a packed pair does not guarantee.
__device__ __forceinline__ void stft_emit(const float2 v, int i, int64_t cplx_base, int64_t mag_base, float2* __restrict__ cplx_out,
                                          const ActView& mag_out) {
  if (cplx_base >= 0) cplx_out[cplx_base + i] = v;
  if (mag_base >= 0) {
    const float a = hypotf(v.x, v.y);         // tf.abs(complex64)
    if (mag_out.fmt == ACT_F32) {
      reinterpret_cast<float*>(mag_out.p)[mag_base + i] = a;
    } else {                                  // split-bf16 planes for the tensor-core encoder
      const __nv_bfloat16 h = __float2bfloat16_rn(a);
      reinterpret_cast<__nv_bfloat16*>(mag_out.p)[mag_base + i] = h;
      if (mag_out.plane != 0)
        reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(mag_out.p) + mag_out.plane)[mag_base + i] =
            __float2bfloat16_rn(a - __bfloat162float(h));
    }
  }
}

// The samples of a CTA's two frames are ONE contiguous span of the 48 kHz stream (wind + hop samples): it is staged into
// shared memory by the TMA unit (cp.async.bulk, one elected thread, mbarrier completion) -- rows of 52799 floats start at
// any multiple of 4 bytes, so the copy starts at the 16-byte boundary below the span -- and windowed from there.
__device__ __forceinline__ uint32_t stft_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) stft_kernel(const float* __restrict__ x, int n_samples, int hop, const FftPlan p,
                                                   int fbase, int fend, int frame0, int n_frames_out, float2* __restrict__ cplx_out,
                                                   int mag0, int n_mag, const ActView mag_out) {
  extern __shared__ __align__(16) float2 smem[];
  __shared__ __align__(8) unsigned long long stage_bar;
  float2* buf0 = smem;
  float2* buf1 = smem + 2 * p.n;
  const int n = p.n;
  const int fa = fbase + 2 * blockIdx.x, fb = fa + 1;
  const bool has_b = fb < fend;
  const int row = blockIdx.y;
  const float* xa = x + (int64_t)row * n_samples + (int64_t)fa * hop;
  // aligned span [xs, xs + span) covering the frames' samples; the whole input is gridDim.y rows of n_samples floats
  const int mis = (int)((reinterpret_cast<uintptr_t>(xa) & 15) >> 2);
  const float* xs = xa - mis;
  const int span = (mis + n + (has_b ? hop : 0) + 3) & ~3;
  const bool staged = xs + span <= x + (int64_t)gridDim.y * n_samples && span <= 4 * n;   // (buf1 holds 4n floats)
  const uint32_t bar = stft_smem_u32(&stage_bar);
  if (staged && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_prologue();
  const float* raw = xa;
  if (staged) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)span * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stft_smem_u32(buf1)),
                   "l"(xs), "r"(bytes), "r"(bar)
                   : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.u32 %0, 1, 0, q;\n}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
    raw = reinterpret_cast<const float*>(buf1) + mis;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float w = __ldg(p.hann + i);
    buf0[i] = make_float2(raw[i] * w, 0.f);
    buf0[n + i] = make_float2(has_b ? raw[hop + i] * w : 0.f, 0.f);
  }
  // (block_fft_nf synchronises before its first pass writes buf1: the staged samples are consumed by then)
  const float2* z = block_fft_nf(buf0, buf1, p, 2);
  // where the two frames go (-1: not wanted)
  auto bases = [&](int f, int64_t& cb, int64_t& mb) {
    cb = (cplx_out != nullptr && f >= frame0 && f < frame0 + n_frames_out) ? ((int64_t)row * n_frames_out + (f - frame0)) * n : -1;
    mb = (mag_out.p != nullptr && f >= mag0 && f < mag0 + n_mag) ? ((int64_t)row * n_mag + (f - mag0)) * n : -1;
  };
  int64_t ca, ma, cb, mb;
  bases(fa, ca, ma);
  bases(fb, cb, mb);
  if (!has_b) { cb = -1; mb = -1; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    stft_emit(z[i], i, ca, ma, cplx_out, mag_out);
    if (has_b) stft_emit(z[n + i], i, cb, mb, cplx_out, mag_out);
  }
}

int launch_stft(const float* x, int rows, int n_samples, int wind, int hop, int n_frames_total, int frame0,
                int n_frames_out, float* cplx_out, int mag0, int n_mag, const ActView& mag_out, cudaStream_t st) {
  SAG_REQUIRE(rows > 0 && wind > 0 && hop > 0, SAG_EINVAL, "stft: bad arguments");
  SAG_REQUIRE((int64_t)(n_frames_total - 1) * hop + wind <= n_samples, SAG_EINVAL,
              "stft: %d frames of %d (hop %d) exceed %d samples", n_frames_total, wind, hop, n_samples);
  int lo = n_frames_total, hi = 0;
  if (cplx_out != nullptr && n_frames_out > 0) { lo = std::min(lo, frame0); hi = std::max(hi, frame0 + n_frames_out); }
  if (mag_out.p != nullptr && n_mag > 0) { lo = std::min(lo, mag0); hi = std::max(hi, mag0 + n_mag); }
  if (hi <= lo) return SAG_OK;
  SAG_REQUIRE(lo >= 0 && hi <= n_frames_total, SAG_EINVAL, "stft: frame range [%d,%d) outside [0,%d)", lo, hi, n_frames_total);
  FftPlan p;
  SAG_TRY(get_plan(wind, &p));
  size_t smem = 4 * sizeof(float2) * wind;
  if (smem > 48 * 1024) SAG_CHECK_CUDA(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((hi - lo + 1) / 2, rows);
  launch_pdl(stft_kernel, grid, dim3(256), smem, st, x, n_samples, hop, p, lo, hi, frame0, n_frames_out, reinterpret_cast<float2*>(cplx_out),
             mag0, n_mag, mag_out);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- K6: (sigmoid mask x STFT) -> real inverse FFT -> overlap-add / n_overlap -> crop ------------------------------
// One CTA per (window, PAIR of tracks).  Only the real part of ifft(m.S) is kept (myutils.py:191-192), and
//   real(ifft(Y)) = ifft(Yh),  Yh[k] = (Y[k] + conj(Y[N-k])) / 2   (Hermitian part),
// so two tracks ride one complex transform: Z = Yh_a + i Yh_b  ->  ifft(Z) = y_a + i y_b.  NFR frames are transformed
// together (one set of block-wide syncs and twiddle fetches for all of them); every mask element is read, and its
// sigmoid evaluated, exactly once.  Frame-space position of output sample j is j + (n_overlap-1)*hop
// (myutils.py:198-205); each thread owns output positions, so the overlap-add needs no atomics.
constexpr int ISTFT_NFR = 4;
constexpr int ISTFT_SLOTS = 10;            // register overlap-add: output positions tid + 256*s, s < 10 -> segments of 2560 samples
constexpr int ISTFT_SEG = 256 * ISTFT_SLOTS;
// The output range is cut into segments of ISTFT_SEG samples (blockIdx.y), each with its own frame range: twice the
// CTAs of half the length fill the SMs' CTA slots evenly (512 long CTAs over 444 slots left the second wave 15 % full),
// and the shorter accumulator arrays free registers for a fourth CTA per SM.
// The overlap-add accumulators live in registers (each thread owns positions tid + 256*s of its segment).
__global__ void __launch_bounds__(256, 4) istft_pair_kernel(const float2* __restrict__ S, const float* __restrict__ mask,
                                                         int apply_sigmoid, int tracks, int n_frames, const FftPlan p,
                                                         int hop, int nf_total, int p0_all, int n_out_all, float inv_scale,
                                                         int nfr, float* __restrict__ out) {
  extern __shared__ __align__(16) float2 smem[];
  const int n = p.n;
  float2* buf0 = smem;
  float2* buf1 = smem + nfr * n;
  float ra[ISTFT_SLOTS], rb[ISTFT_SLOTS];
  pdl_prologue();
  // this CTA's segment of the output and the frames that reach it
  const int seg0 = (int)blockIdx.y * ISTFT_SEG;
  const int n_out = min(ISTFT_SEG, n_out_all - seg0);
  const int p0 = p0_all + seg0;
  int f_lo = p0 - n + 1 <= 0 ? 0 : (p0 - n + 1 + hop - 1) / hop;
  int f_hi = min((p0 + n_out - 1) / hop, nf_total - 1);
  const int pairs = (tracks + 1) / 2;
  const int64_t row = blockIdx.x / pairs;
  const int ka = (int)(blockIdx.x % pairs) * 2, kb = ka + 1;
  const bool has_b = kb < tracks;
#pragma unroll
  for (int sl = 0; sl < ISTFT_SLOTS; ++sl) { ra[sl] = 0.f; rb[sl] = 0.f; }
  const float* ma = mask != nullptr ? mask + (row * tracks + ka) * (int64_t)n_frames * n : nullptr;
  const float* mb = (mask != nullptr && has_b) ? mask + (row * tracks + kb) * (int64_t)n_frames * n : nullptr;
  for (int f0 = f_lo; f0 <= f_hi; f0 += nfr) {
    const int nf = min(nfr, f_hi - f0 + 1);
    __syncthreads();                                   // previous group's transforms fully consumed
    for (int g = 0; g < nf; ++g) {
      const float2* s = S + (row * n_frames + (f0 + g)) * (int64_t)n;
      const int64_t mo = (int64_t)(f0 + g) * n;
      constexpr int U = 3;                               // bins per thread per batch: all loads first, then the math
      for (int kb0 = 0; kb0 <= n / 2; kb0 += U * 256) {
        float2 xk[U], xn[U];
        float mak[U], man[U], mbk[U], mbn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k = kb0 + threadIdx.x + 256 * u;
          xk[u] = xn[u] = make_float2(0.f, 0.f);
          mak[u] = man[u] = mbk[u] = mbn[u] = 0.f;
          if (k <= n / 2) {
            const int kn = k == 0 ? 0 : n - k;
            xk[u] = __ldg(s + k); xn[u] = __ldg(s + kn);
            if (ma != nullptr) { mak[u] = __ldg(ma + mo + k); man[u] = __ldg(ma + mo + kn); }
            if (mb != nullptr) { mbk[u] = __ldg(mb + mo + k); mbn[u] = __ldg(mb + mo + kn); }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k = kb0 + threadIdx.x + 256 * u;
          if (k > n / 2) continue;
          const int kn = k == 0 ? 0 : n - k;
          float gak = 1.f, gan = 1.f, gbk = has_b ? 1.f : 0.f, gbn = gbk;
          if (ma != nullptr) {
            gak = mak[u]; gan = man[u];
            if (apply_sigmoid) { gak = 1.f / (1.f + expf(-gak)); gan = 1.f / (1.f + expf(-gan)); }
          }
          if (mb != nullptr) {
            gbk = mbk[u]; gbn = mbn[u];
            if (apply_sigmoid) { gbk = 1.f / (1.f + expf(-gbk)); gbn = 1.f / (1.f + expf(-gbn)); }
          }
          // Hermitian parts of the two masked spectra at bin k
          const float2 ya = make_float2(0.5f * (gak * xk[u].x + gan * xn[u].x), 0.5f * (gak * xk[u].y - gan * xn[u].y));
          const float2 yb = make_float2(0.5f * (gbk * xk[u].x + gbn * xn[u].x), 0.5f * (gbk * xk[u].y - gbn * xn[u].y));
          // Z[k] = ya + i yb ; Z[N-k] = conj(ya) + i conj(yb); stored conjugated: ifft(Z) = conj(fft(conj Z)) / N
          buf0[g * n + k] = make_float2(ya.x - yb.y, -(ya.y + yb.x));
          buf0[g * n + kn] = make_float2(ya.x + yb.y, -(yb.x - ya.y));
        }
      }
    }
    const float2* res = block_fft_nf(buf0, buf1, p, nf);     // res = N * (y_a - i y_b)
    const int lo = max(f0 * hop - p0, 0), hi = min((f0 + nf - 1) * hop - p0 + n, n_out);
#pragma unroll
    for (int sl = 0; sl < ISTFT_SLOTS; ++sl) {
      const int j = threadIdx.x + 256 * sl;
      if (j >= lo && j < hi) {
        float aa = 0.f, ab = 0.f;
        for (int g = 0; g < nf; ++g) {
          const int i = j - ((f0 + g) * hop - p0);
          if (i >= 0 && i < n) { const float2 r = res[g * n + i]; aa += r.x; ab -= r.y; }
        }
        ra[sl] += aa;
        rb[sl] += ab;
      }
    }
  }
  float* oa = out + (row * tracks + ka) * (int64_t)n_out_all + seg0;
  float* ob = out + (row * tracks + kb) * (int64_t)n_out_all + seg0;
#pragma unroll
  for (int sl = 0; sl < ISTFT_SLOTS; ++sl) {
    const int j = threadIdx.x + 256 * sl;
    if (j < n_out) {
      oa[j] = ra[sl] * inv_scale;
      if (has_b) ob[j] = rb[sl] * inv_scale;
    }
  }
}

int launch_istft(const float* S, const float* mask, int apply_sigmoid, int rows_s, int tracks, int n_frames, int wind,
                 int n_overlap, int crop0, int n_out, float* out, cudaStream_t st) {
  SAG_REQUIRE(rows_s > 0 && tracks > 0 && n_overlap > 0 && wind % n_overlap == 0, SAG_EINVAL, "istft: bad arguments");
  const int hop = wind / n_overlap;
  const int nf = (n_frames / n_overlap) * n_overlap;      // myutils.py:187-188
  const int full = (nf / n_overlap) * wind - (n_overlap - 1) * hop;
  SAG_REQUIRE(nf > 0 && crop0 >= 0 && n_out > 0 && crop0 + n_out <= full, SAG_EINVAL,
              "istft: crop [%d,%d) outside the %d output samples", crop0, crop0 + n_out, full);
  FftPlan p;
  SAG_TRY(get_plan(wind, &p));
  const int p0 = crop0 + (n_overlap - 1) * hop;
  const int n_seg = cdiv(n_out, ISTFT_SEG);
  int nfr = ISTFT_NFR;
  size_t smem = 0;
  static const int smem_cap_kb = [] { const char* v = getenv("SAG_ISTFT_SMEM_KB"); return v ? atoi(v) : 56; }();   // tuning knob
  for (; nfr >= 1; nfr >>= 1) {
    smem = 2 * sizeof(float2) * (size_t)wind * nfr;
    if (smem <= (size_t)smem_cap_kb * 1024 || nfr == 1) break;   // four CTAs per SM when possible
  }
  SAG_REQUIRE(smem <= 220 * 1024, SAG_EUNSUPPORTED, "istft: %zu bytes of shared memory needed", smem);
  const float inv_scale = 1.0f / ((float)wind * (float)n_overlap);
  const int pairs = (tracks + 1) / 2;
  SAG_CHECK_CUDA(cudaFuncSetAttribute(istft_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_pdl(istft_pair_kernel, dim3(rows_s * pairs, n_seg), dim3(256), smem, st, reinterpret_cast<const float2*>(S), mask, apply_sigmoid,
             tracks, n_frames, p, hop, nf, p0, n_out, inv_scale, nfr, out);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- K6+K9 fused: masked inverse STFT + time-varying mixing -----------------------------------------------------------
// The decode of inference_ops (model.py:424-432) is linear in the separated tracks, and so are ifft / overlap-add:
//   out[n, o] = sum_k W[o, k, seg(n)] * x_sep[k, n] + b[o, seg(n)]
//             = OLA( real ifft( S[f] * G[o, seg][f] ) )[n] / (n_overlap * N) + b[o, seg],
//   G[o, seg][f, bin] = sum_k W[o, k, seg] * sigmoid(m_k[f, bin])
// for the samples n of one localization segment `seg` (the weights are piecewise constant: model.py:262-263).  So instead
// of 32 inverse transforms per frame (16 with two tracks per complex FFT) the 3 output channels need 1.5.
//   mask_gains_kernel : streams the 32 mask planes once (one thread per (window, frame, bin), 32 independent coalesced
//                       loads in flight) and folds them into the 9 gains (3 segments x 3 channels) -- HBM/L2 bound.
//   istft_mix_kernel  : one CTA per (window, segment): channels (Y, Z) ride one complex transform per frame, channel X of
//                       two consecutive frames the third; register overlap-add; + bias; writes (B, T, 3).
// x_sep is never formed: this is the production path (sag_forward); inference_ops, which exposes `sep_channels`, keeps
// the two-kernel path (istft_pair_kernel + mix_kernel).
constexpr int MIX_SLOTS = 8;              // output positions tid + 256*s of a segment (segments up to 2048 samples)
constexpr int GAIN_MAX_TRACKS = 64;

// gains (rows, segments*3, n_frames, n) <- mask (rows*tracks, n_frames, n) logits, loc (rows, segments, 3*(tracks+1)).
// One thread per 4 consecutive bins: 16-byte loads, every weight fetched from shared memory serves 4 bins.
__global__ void __launch_bounds__(256) mask_gains_kernel(const float* __restrict__ mask, const float* __restrict__ loc, int tracks,
                                                         int n_frames, int n, int segments, float* __restrict__ gains) {
  extern __shared__ float s_w[];          // [tracks][12]: the (up to) 9 weights of a track side by side
  pdl_prologue();
  const int per_frame = n / 1024;
  const int b = blockIdx.y, f = blockIdx.x / per_frame, k = ((blockIdx.x % per_frame) * 256 + threadIdx.x) * 4;
  const int K1 = tracks + 1, ng = segments * 3;
  for (int i = threadIdx.x; i < 12 * tracks; i += blockDim.x) {
    const int kk = i / 12, gi = i % 12;
    s_w[i] = gi < ng ? __ldg(loc + ((int64_t)b * ng + gi) * K1 + kk) : 0.f;
  }
  __syncthreads();
  const int64_t plane = (int64_t)n_frames * n;
  const float* mp = mask + (int64_t)b * tracks * plane + (int64_t)f * n + k;
  float4 acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = 0; k0 < tracks; k0 += 8) {
    float4 m[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      m[u] = k0 + u < tracks ? __ldg(reinterpret_cast<const float4*>(mp + (int64_t)(k0 + u) * plane)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (k0 + u >= tracks) break;
      float4 sg;                                             // tf.sigmoid (model.py:334)
      sg.x = sigmoid_sfu(m[u].x); sg.y = sigmoid_sfu(m[u].y);
      sg.z = sigmoid_sfu(m[u].z); sg.w = sigmoid_sfu(m[u].w);
      const float4* w4 = reinterpret_cast<const float4*>(s_w + (k0 + u) * 12);
      const float4 wa = w4[0], wb = w4[1], wc = w4[2];
      const float w[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        acc[i].x = fmaf(w[i], sg.x, acc[i].x); acc[i].y = fmaf(w[i], sg.y, acc[i].y);
        acc[i].z = fmaf(w[i], sg.z, acc[i].z); acc[i].w = fmaf(w[i], sg.w, acc[i].w);
      }
    }
  }
  float* gp = gains + (int64_t)b * ng * plane + (int64_t)f * n + k;
#pragma unroll
  for (int i = 0; i < 9; ++i)
    if (i < ng) *reinterpret_cast<float4*>(gp + (int64_t)i * plane) = acc[i];
}

__global__ void __launch_bounds__(256, 2) istft_mix_kernel(const float2* __restrict__ S, const float* __restrict__ gains,
                                                           const float* __restrict__ loc, int tracks, int n_frames,
                                                           const FftPlan p, int hop, int nf_total, int p0_all, int t_out,
                                                           int segments, int n_sub, float inv_scale, float* __restrict__ out) {
  extern __shared__ __align__(16) float2 smem[];
  const int n = p.n;
  float2* buf0 = smem;                    // [3][n]: frame f (Y, Z), frame f+1 (Y, Z), X of both frames
  float2* buf1 = smem + 3 * n;
  pdl_prologue();
  // CTA -> (window, localization segment, part of the segment): parts shorten the frame loop and fill more SMs
  const int sub = blockIdx.x % n_sub, seg = (blockIdx.x / n_sub) % segments, b = blockIdx.x / (n_sub * segments);
  const int seg_len = t_out / segments / n_sub, K1 = tracks + 1;
  const int out0 = seg * (t_out / segments) + sub * seg_len;  // first output sample of this CTA
  const int p0 = p0_all + out0;                              // its frame-space position
  const int f_lo = p0 - n + 1 <= 0 ? 0 : (p0 - n + 1 + hop - 1) / hop;
  const int f_hi = min((p0 + seg_len - 1) / hop, nf_total - 1);
  float acc[3][MIX_SLOTS];
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int sl = 0; sl < MIX_SLOTS; ++sl) acc[o][sl] = 0.f;
  const int64_t plane = (int64_t)n_frames * n;               // between gain planes
  const float* gb = gains + ((int64_t)b * segments + seg) * 3 * plane;
  for (int f0 = f_lo; f0 <= f_hi; f0 += 2) {
    const int nf = min(2, f_hi - f0 + 1);
    __syncthreads();                                         // previous pair fully consumed
    if (nf == 1)
      for (int i = threadIdx.x; i < n; i += blockDim.x) buf0[n + i] = make_float2(0.f, 0.f);
    for (int g = 0; g < nf; ++g) {
      const float2* s = S + ((int64_t)b * n_frames + (f0 + g)) * (int64_t)n;
      const float* gf = gb + (int64_t)(f0 + g) * n;
      for (int k = threadIdx.x; k <= n / 2; k += blockDim.x) {
        const int kn = k == 0 ? 0 : n - k;
        const float2 xk = __ldg(s + k), xn = __ldg(s + kn);
        float gk[3], gn[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) { gk[o] = __ldg(gf + o * plane + k); gn[o] = __ldg(gf + o * plane + kn); }
        // Hermitian parts of the three gain-weighted spectra at bin k (see istft_pair_kernel)
        float2 y[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) y[o] = make_float2(0.5f * (gk[o] * xk.x + gn[o] * xn.x), 0.5f * (gk[o] * xk.y - gn[o] * xn.y));
        // Z = y0 + i y1, stored conjugated: ifft(Z) = conj(fft(conj Z)) / N
        buf0[g * n + k] = make_float2(y[0].x - y[1].y, -(y[0].y + y[1].x));
        buf0[g * n + kn] = make_float2(y[0].x + y[1].y, -(y[1].x - y[0].y));
        // third transform: X of frame f0 as the real carrier, X of frame f0+1 as the imaginary one (same thread owns the
        // bin in both rounds)
        if (g == 0) {
          buf0[2 * n + k] = make_float2(y[2].x, -y[2].y);
          buf0[2 * n + kn] = make_float2(y[2].x, y[2].y);
        } else {
          float2 u = buf0[2 * n + k];
          buf0[2 * n + k] = make_float2(u.x - y[2].y, u.y - y[2].x);
          if (kn != k) {
            u = buf0[2 * n + kn];
            buf0[2 * n + kn] = make_float2(u.x + y[2].y, u.y - y[2].x);
          }
        }
      }
    }
    const float2* res = block_fft_nf(buf0, buf1, p, 3);       // res[0..1] = N (y_Y - i y_Z) of the two frames, res[2] = N (x_f - i x_f+1)
#pragma unroll
    for (int sl = 0; sl < MIX_SLOTS; ++sl) {
      const int j = threadIdx.x + 256 * sl;
      if (j < seg_len) {
        const int i0 = p0 + j - f0 * hop, i1 = i0 - hop;       // offsets inside frame f0 / f0+1
        if (i0 >= 0 && i0 < n) {
          const float2 r = res[i0], x = res[2 * n + i0];
          acc[0][sl] += r.x; acc[1][sl] -= r.y; acc[2][sl] += x.x;
        }
        if (nf == 2 && i1 >= 0 && i1 < n) {
          const float2 r = res[n + i1], x = res[2 * n + i1];
          acc[0][sl] += r.x; acc[1][sl] -= r.y; acc[2][sl] -= x.y;
        }
      }
    }
  }
  const float* lb = loc + ((int64_t)b * segments + seg) * 3 * K1 + tracks;     // bias of channel o at lb[o * K1]
  float* ob = out + ((int64_t)b * t_out + out0) * 3;
#pragma unroll
  for (int sl = 0; sl < MIX_SLOTS; ++sl) {
    const int j = threadIdx.x + 256 * sl;
    if (j < seg_len) {
#pragma unroll
      for (int o = 0; o < 3; ++o) ob[(int64_t)j * 3 + o] = fmaf(acc[o][sl], inv_scale, __ldg(lb + o * K1));
    }
  }
}

int istft_mix_supported(int tracks, int t_out, int segments, int wind) {
  return segments > 0 && segments <= 3 && t_out % segments == 0 && t_out / segments <= 256 * MIX_SLOTS && tracks >= 1 &&
         tracks <= GAIN_MAX_TRACKS && wind % 1024 == 0;
}
size_t istft_mix_gain_floats(int rows, int n_frames, int wind, int segments) { return (size_t)rows * segments * 3 * n_frames * wind; }

// S (rows, n_frames, wind) complex, mask (rows*tracks, n_frames, wind) logits, loc (rows, segments, 3*(tracks+1));
// gains: scratch of istft_mix_gain_floats() floats; out (rows, t_out, 3) = samples [crop0, crop0 + t_out) of the mixed
// inverse STFT.  mask == null: `gains` already holds the folded masks (deconv1's fused epilogue wrote them).
int launch_istft_mix(const float* S, const float* mask, const float* loc, float* gains, int rows, int tracks, int n_frames, int wind,
                     int n_overlap, int crop0, int t_out, int segments, float* out, cudaStream_t st) {
  SAG_REQUIRE(rows > 0 && n_overlap > 0 && wind % n_overlap == 0 && loc != nullptr && gains != nullptr, SAG_EINVAL, "istft_mix: bad arguments");
  SAG_REQUIRE(istft_mix_supported(tracks, t_out, segments, wind), SAG_EUNSUPPORTED, "istft_mix: %d samples / %d segments / %d tracks", t_out, segments, tracks);
  const int hop = wind / n_overlap;
  const int nf = (n_frames / n_overlap) * n_overlap;      // myutils.py:187-188
  const int full = (nf / n_overlap) * wind - (n_overlap - 1) * hop;
  SAG_REQUIRE(nf > 0 && crop0 >= 0 && crop0 + t_out <= full, SAG_EINVAL, "istft_mix: crop [%d,%d) outside the %d output samples", crop0, crop0 + t_out, full);
  FftPlan p;
  SAG_TRY(get_plan(wind, &p));
  if (mask != nullptr) {
    launch_pdl(mask_gains_kernel, dim3(n_frames * (wind / 1024), rows), dim3(256), sizeof(float) * 12 * tracks, st, mask, loc, tracks,
               n_frames, wind, segments, gains);
    SAG_LAUNCH_CHECK();
  }
  const size_t smem = 6 * sizeof(float2) * (size_t)wind;
  SAG_REQUIRE(smem <= 220 * 1024, SAG_EUNSUPPORTED, "istft_mix: %zu bytes of shared memory needed", smem);
  SAG_CHECK_CUDA(cudaFuncSetAttribute(istft_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const int sub_env = [] { const char* v = getenv("SAG_ISTFT_MIX_SUB"); return v ? atoi(v) : 2; }();   // tuning knob
  const int seg_len = t_out / segments;
  const int n_sub = (sub_env >= 1 && seg_len % sub_env == 0) ? sub_env : 1;
  launch_pdl(istft_mix_kernel, dim3(rows * segments * n_sub), dim3(256), smem, st, reinterpret_cast<const float2*>(S), (const float*)gains, loc,
             tracks, n_frames, p, hop, nf, crop0 + (n_overlap - 1) * hop, t_out, segments, n_sub, 1.0f / ((float)wind * (float)n_overlap), out);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// frames [f_lo, f_hi] of the inverse STFT that reach output samples [crop0, crop0+n_out)
void istft_needed_frames(int n_frames, int wind, int n_overlap, int crop0, int n_out, int* f_lo, int* f_hi) {
  const int hop = wind / n_overlap;
  const int nf = (n_frames / n_overlap) * n_overlap;
  const int p0 = crop0 + (n_overlap - 1) * hop, p1 = p0 + n_out;
  int lo = (p0 - wind + 1 + hop - 1) / hop;
  if (p0 - wind + 1 <= 0) lo = 0;
  int hi = (p1 - 1) / hop;
  if (hi > nf - 1) hi = nf - 1;
  *f_lo = lo;
  *f_hi = hi;
}

}  // namespace sag
