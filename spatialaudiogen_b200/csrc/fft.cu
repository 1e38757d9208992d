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
// grid (n_frames_launch, rows); frame index = fbase + blockIdx.x.
__global__ void __launch_bounds__(256) stft_kernel(const float* __restrict__ x, int n_samples, int hop, const FftPlan p,
                                                   int fbase, int frame0, int n_frames_out, float2* __restrict__ cplx_out,
                                                   int mag0, int n_mag, const ActView mag_out) {
  extern __shared__ __align__(16) float2 smem[];
  float2* buf0 = smem;
  float2* buf1 = smem + p.n;
  pdl_prologue();
  const int f = fbase + blockIdx.x;
  const int row = blockIdx.y;
  const float* xr = x + (int64_t)row * n_samples + (int64_t)f * hop;
  for (int i = threadIdx.x; i < p.n; i += blockDim.x) buf0[i] = make_float2(__ldg(xr + i) * __ldg(p.hann + i), 0.f);
  float2* res = block_fft(buf0, buf1, p);
  if (cplx_out != nullptr && f >= frame0 && f < frame0 + n_frames_out) {
    float2* o = cplx_out + ((int64_t)row * n_frames_out + (f - frame0)) * p.n;
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) o[i] = res[i];
  }
  if (mag_out.p != nullptr && f >= mag0 && f < mag0 + n_mag) {
    const int64_t e0 = ((int64_t)row * n_mag + (f - mag0)) * p.n;
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
      float2 v = res[i];
      const float a = hypotf(v.x, v.y);       // tf.abs(complex64)
      if (mag_out.fmt == ACT_F32) {
        reinterpret_cast<float*>(mag_out.p)[e0 + i] = a;
      } else {                                // split-bf16 planes for the tensor-core encoder
        const __nv_bfloat16 h = __float2bfloat16_rn(a);
        reinterpret_cast<__nv_bfloat16*>(mag_out.p)[e0 + i] = h;
        if (mag_out.plane != 0)
          reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(mag_out.p) + mag_out.plane)[e0 + i] =
              __float2bfloat16_rn(a - __bfloat162float(h));
      }
    }
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
  size_t smem = 2 * sizeof(float2) * wind;
  if (smem > 48 * 1024) SAG_CHECK_CUDA(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(hi - lo, rows);
  launch_pdl(stft_kernel, grid, dim3(256), smem, st, x, n_samples, hop, p, lo, frame0, n_frames_out, reinterpret_cast<float2*>(cplx_out),
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
  for (; nfr >= 1; nfr >>= 1) {
    smem = 2 * sizeof(float2) * (size_t)wind * nfr;
    if (smem <= 56 * 1024 || nfr == 1) break;              // four CTAs per SM when possible
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
