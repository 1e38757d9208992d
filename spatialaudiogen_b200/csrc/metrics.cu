// Evaluation metrics of the reference as GPU kernels.
//   metrics_kernel : evaluation_ops per-sample outputs (reference model.py:110-154): STFT distance
//                    (_stft_mse_ops :62-76 + myutils.stft_for_loss myutils.py:151-178), LSD (_lsd_ops :78-94 +
//                    myutils.stft with window 1200, overlap 2), temporal MSE (:96-99), SNR (:101-108); the Hilbert
//                    envelope distance of myutils.compute_envelope_dist (myutils.py:109-116) and the amplitudes of
//                    eval.py:197-198.  One CTA per (window, channel); all FFTs run in shared memory.
//   sh_rms_kernel  : AmbiDecoder.decode 'projection' (decoder.py:24-26) on the spherical mesh of distance.py:9-13
//                    followed by the RMS map of distance.py:48-51, as a warp-shuffle reduction over time.
// HBM-bound (each input sample is read once); algorithmic bytes = 2*T*3*4 per window for the metrics,
// T*4*4 per window for the RMS map.
#include "common.cuh"
#include "fft_device.cuh"
#include <mutex>
#include <cmath>

namespace sag {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the block; result valid in every thread. red: >= 32 floats of shared memory.
__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}
__device__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
  if (wid == 0) {
    t = warp_max(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

struct MetricPlans {
  FftPlan stft;   // 2048 (ceil-pow2 of int(0.025*rate), myutils.py:155)
  FftPlan lsd;    // 1200 (int(0.025*rate), model.py:123)
  FftPlan env;    // T (scipy.signal.hilbert over the whole 0.1 s window)
};

// Hann-windowed frames of gt and pred at sample offset `off`, transformed together (one set of block-wide passes):
// result[0..n) = FFT(gt frame), result[n..2n) = FFT(pred frame)
__device__ const float2* windowed_fft_pair(const float* sg, const float* sp, int off, const FftPlan& p, float2* b0, float2* b1) {
  __syncthreads();
  for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
    const float w = __ldg(p.hann + i);
    b0[i] = make_float2(sg[off + i] * w, 0.f);
    b0[p.n + i] = make_float2(sp[off + i] * w, 0.f);
  }
  return block_fft_nf(b0, b1, p, 2);
}

// One CTA per (task, window, channel); tasks: 0 = Hilbert envelope distance (the longest, scheduled first),
// 1 = STFT distance + temporal MSE / SNR / amplitudes, 2 = LSD.  gt and pred ride every transform together.
__global__ void __launch_bounds__(1024) metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int batch,
                                                      int t, const MetricPlans P, int has_env, float* __restrict__ stft_ps,
                                                      float* __restrict__ lsd_ps, float* __restrict__ mse_ps,
                                                      float* __restrict__ snr_ps, float* __restrict__ env_ps,
                                                      float* __restrict__ amp) {
  extern __shared__ __align__(16) float2 smem[];
  __shared__ float red[32];
  const int nwc = batch * 3;
  int task = blockIdx.x / nwc;
  const int wc = blockIdx.x % nwc;                     // window * 3 + channel
  if (!has_env) task += 1;                             // grid holds only tasks 1 and 2
  const int b = wc / 3, ch = wc % 3;
  const int nmax = task == 0 ? P.env.n : (task == 1 ? P.stft.n : P.lsd.n);
  float2* b0 = smem;
  float2* b1 = smem + 2 * nmax;
  float* sg = reinterpret_cast<float*>(smem + 4 * nmax);   // gt signal [t]
  float* sp = sg + t;                                      // pred signal [t]

  const float* pg = gt + (int64_t)b * t * 3 + ch;
  const float* pp = pred + (int64_t)b * t * 3 + ch;
  float se = 0.f, sgg = 0.f, mp = 0.f, mg = 0.f;
  for (int i = threadIdx.x; i < t; i += blockDim.x) {
    float g = __ldg(pg + (int64_t)i * 3), p = __ldg(pp + (int64_t)i * 3);
    sg[i] = g;
    sp[i] = p;
    float d = g - p;
    se = fmaf(d, d, se);
    sgg = fmaf(g, g, sgg);
    mp = fmaxf(mp, fabsf(p));
    mg = fmaxf(mg, fabsf(g));
  }

  if (task == 1) {
    se = block_sum(se, red);
    sgg = block_sum(sgg, red);
    mp = block_max(mp, red);
    mg = block_max(mg, red);
    if (threadIdx.x == 0) {
      mse_ps[wc] = se / (float)t;                                         // model.py:99
      snr_ps[wc] = 10.f * logf((sgg + 0.1f) / (se + 0.1f)) / logf(10.f);  // model.py:105-107
      atomicMax(reinterpret_cast<int*>(amp + b * 2 + 0), __float_as_int(mp));      // eval.py:197-198 (values >= 0)
      atomicMax(reinterpret_cast<int*>(amp + b * 2 + 1), __float_as_int(mg));
    }
    // ---- STFT distance: windows of n=2048 at stride n/2 grouped as myutils.py:167-173 builds them ----
    const int n = P.stft.n, stride = n / 2;
    float acc_total = 0.f;
    int nwin_total = 0;
    for (int i = 0; i < 2; ++i) {
      int nW = (t - i * stride - 1) / n;
      for (int wdx = 0; wdx < nW; ++wdx) {
        const float2* r = windowed_fft_pair(sg, sp, i * stride + wdx * n, P.stft, b0, b1);
        float a = 0.f;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
          float dx = r[k].x - r[n + k].x, dy = r[k].y - r[n + k].y;
          a += dx * dx + dy * dy;                                          // |stft_gt - stft_pred|^2
        }
        acc_total += block_sum(a, red) / (float)n;                         // mean over freq
        ++nwin_total;
      }
    }
    if (threadIdx.x == 0) stft_ps[wc] = nwin_total > 0 ? acc_total / (float)nwin_total : 0.f;
  } else if (task == 2) {
    // ---- LSD: myutils.stft(x, 1200, 2): n_winds = t/1200 - 1, frames at hop 600 ----
    const int n = P.lsd.n, hop = n / 2;
    const int nframes = 2 * (t / n - 1);
    float acc_total = 0.f;
    const float k10 = 10.f / logf(10.f);
    for (int f = 0; f < nframes; ++f) {
      const float2* r = windowed_fft_pair(sg, sp, f * hop, P.lsd, b0, b1);
      float a = 0.f;
      for (int k = threadIdx.x; k < n; k += blockDim.x) {
        float d = k10 * (logf(hypotf(r[k].x, r[k].y) + 1e-2f) - logf(hypotf(r[n + k].x, r[n + k].y) + 1e-2f));
        a = fmaf(d, d, a);
      }
      acc_total += sqrtf(block_sum(a, red) / (float)n);
    }
    if (threadIdx.x == 0) lsd_ps[wc] = nframes > 0 ? acc_total / (float)nframes : 0.f;
  } else {
    // ---- Hilbert envelope distance: analytic signals via FFT (scipy.signal.hilbert), N = t, gt and pred together ----
    const int n = P.env.n;
    const float inv_n = 1.f / (float)n;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      b0[i] = make_float2(sg[i], 0.f);
      b0[n + i] = make_float2(sp[i], 0.f);
    }
    float2* r = block_fft_nf(b0, b1, P.env, 2);
    float2* o = (r == b0) ? b1 : b0;
    // h[0]=1, h[1..n/2-1]=2, h[n/2]=1 (n even), rest 0 ; then inverse FFT via conj trick
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      float h;
      if ((n & 1) == 0) h = (k == 0 || k == n / 2) ? 1.f : (k < n / 2 ? 2.f : 0.f);
      else h = (k == 0) ? 1.f : (k < (n + 1) / 2 ? 2.f : 0.f);
      o[k] = make_float2(r[k].x * h, -r[k].y * h);
      o[n + k] = make_float2(r[n + k].x * h, -r[n + k].y * h);
    }
    const float2* z = block_fft_nf(o, r, P.env, 2);     // conj(analytic)*n for gt and pred
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float d = (hypotf(z[i].x, z[i].y) - hypotf(z[n + i].x, z[n + i].y)) * inv_n;
      a = fmaf(d, d, a);
    }
    a = block_sum(a, red);
    if (threadIdx.x == 0) env_ps[wc] = sqrtf(a * inv_n);
  }
}

static size_t metrics_smem_bytes(int t, int nmax) { return (size_t)4 * nmax * sizeof(float2) + (size_t)2 * t * sizeof(float); }

int launch_metrics(const float* pred, const float* gt, int batch, int t, int audio_rate, float* stft_ps, float* lsd_ps,
                   float* mse_ps, float* snr_ps, float* env_ps, float* amp, void* scratch, cudaStream_t st) {
  (void)scratch;
  SAG_REQUIRE(batch > 0 && t > 0, SAG_EINVAL, "metrics: bad arguments");
  const int window = (int)(0.025 * (double)audio_rate);                      // definitions.py:10, model.py:123
  int n_stft = 1;
  while (n_stft < window) n_stft <<= 1;                                       // myutils.py:155
  SAG_REQUIRE(t >= 2 * window && t > n_stft, SAG_EINVAL, "metrics: %d samples too short for window %d", t, window);
  MetricPlans P;
  SAG_TRY(get_plan(n_stft, &P.stft));
  SAG_TRY(get_plan(window, &P.lsd));
  SAG_TRY(get_plan(t, &P.env));
  const int has_env = env_ps != nullptr ? 1 : 0;
  const int nmax = std::max(std::max(n_stft, window), has_env ? t : 0);
  size_t smem = metrics_smem_bytes(t, nmax);
  SAG_REQUIRE(smem <= 220 * 1024, SAG_EUNSUPPORTED, "metrics: %zu bytes of shared memory needed", smem);
  SAG_CHECK_CUDA(cudaFuncSetAttribute(metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SAG_CHECK_CUDA(cudaMemsetAsync(amp, 0, sizeof(float) * 2 * batch, st));
  // one CTA per SM (the 4800-point transforms fill the shared memory): wide CTAs hide the latency of the block-wide passes
  // (512 threads: +0.8 % end to end over 256; 1024 no better)
  metrics_kernel<<<batch * 3 * (has_env ? 3 : 2), 512, smem, st>>>(pred, gt, batch, t, P, has_env, stft_ps, lsd_ps, mse_ps,
                                                                  snr_ps, env_ps, amp);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// ---- mel log-spectral distance (reference myutils.compute_lsd_dist, myutils.py:96-106) -----------------------------------
// librosa.feature.melspectrogram(y, sr, n_mels=128, fmax=12000) of librosa 0.6.0 (absent third-party code, restated from
// its published algorithm): centred STFT (reflect padding n_fft/2, periodic Hann, n_fft 2048, hop 512), power spectrum,
// Slaney mel filter bank with area normalisation; then 10*log10(|mel| + 0.01) and the RMS difference over all
// (band, frame) cells.  One CTA per (window, channel); gt and pred ride every transform together.
struct MelBank {
  int n_fft, hop, n_mels, nnz;
  const int* start;      // [n_mels] first FFT bin with a non-zero weight
  const int* len;        // [n_mels]
  const int* off;        // [n_mels] offset into w
  const float* w;        // packed triangle weights
};
static std::mutex g_mel_mu;
static std::map<std::pair<int, int>, MelBank> g_mels;   // (device, rate) -> bank

static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static int get_mel_bank(int rate, MelBank* out) {
  int dev = 0;
  SAG_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mel_mu);
  auto it = g_mels.find({dev, rate});
  if (it != g_mels.end()) { *out = it->second; return SAG_OK; }
  MelBank m;
  m.n_fft = 2048; m.hop = 512; m.n_mels = 128;
  const double fmax = 12000.0, fmin = 0.0;
  const int nb = 1 + m.n_fft / 2;
  std::vector<double> mel_f(m.n_mels + 2);
  const double m0 = hz_to_mel(fmin), m1 = hz_to_mel(fmax);
  for (int i = 0; i < m.n_mels + 2; ++i) mel_f[i] = mel_to_hz(m0 + (m1 - m0) * (double)i / (double)(m.n_mels + 1));
  std::vector<int> start(m.n_mels), len(m.n_mels), off(m.n_mels);
  std::vector<float> w;
  for (int i = 0; i < m.n_mels; ++i) {
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    int first = -1, last = -1;
    std::vector<float> row;
    for (int k = 0; k < nb; ++k) {
      const double f = (double)rate / 2.0 * (double)k / (double)(nb - 1);         // np.linspace(0, sr/2, 1 + n_fft//2)
      const double lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i]), upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1]);
      const double v = std::max(0.0, std::min(lower, upper)) * enorm;
      if (v > 0.0) {
        if (first < 0) first = k;
        last = k;
      }
    }
    start[i] = first < 0 ? 0 : first;
    len[i] = first < 0 ? 0 : last - first + 1;
    off[i] = (int)w.size();
    for (int k = start[i]; k < start[i] + len[i]; ++k) {
      const double f = (double)rate / 2.0 * (double)k / (double)(nb - 1);
      const double lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i]), upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1]);
      w.push_back((float)(std::max(0.0, std::min(lower, upper)) * enorm));
    }
  }
  m.nnz = (int)w.size();
  int *ds = nullptr, *dl = nullptr, *dof = nullptr;
  float* dw = nullptr;
  SAG_CHECK_CUDA(cudaMalloc(&ds, sizeof(int) * m.n_mels));
  SAG_CHECK_CUDA(cudaMalloc(&dl, sizeof(int) * m.n_mels));
  SAG_CHECK_CUDA(cudaMalloc(&dof, sizeof(int) * m.n_mels));
  SAG_CHECK_CUDA(cudaMalloc(&dw, sizeof(float) * std::max<size_t>(w.size(), 1)));
  SAG_CHECK_CUDA(cudaMemcpy(ds, start.data(), sizeof(int) * m.n_mels, cudaMemcpyHostToDevice));
  SAG_CHECK_CUDA(cudaMemcpy(dl, len.data(), sizeof(int) * m.n_mels, cudaMemcpyHostToDevice));
  SAG_CHECK_CUDA(cudaMemcpy(dof, off.data(), sizeof(int) * m.n_mels, cudaMemcpyHostToDevice));
  SAG_CHECK_CUDA(cudaMemcpy(dw, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
  m.start = ds; m.len = dl; m.off = dof; m.w = dw;
  g_mels[{dev, rate}] = m;
  *out = m;
  return SAG_OK;
}

__global__ void __launch_bounds__(256) mel_lsd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int t,
                                                      const FftPlan P, const MelBank mel, int n_frames,
                                                      float* __restrict__ out) {
  extern __shared__ __align__(16) float2 smem[];
  __shared__ float red[32];
  const int n = P.n, nb = n / 2 + 1;
  float2* b0 = smem;
  float2* b1 = smem + 2 * n;
  float* sg = reinterpret_cast<float*>(smem + 4 * n);
  float* sp = sg + t;
  float* pw = sp + t;                                   // [2][nb] power spectra of the current frame
  const int wc = blockIdx.x;                            // window * 3 + channel
  const float* pg = gt + (int64_t)(wc / 3) * t * 3 + wc % 3;
  const float* pp = pred + (int64_t)(wc / 3) * t * 3 + wc % 3;
  for (int i = threadIdx.x; i < t; i += blockDim.x) { sg[i] = __ldg(pg + (int64_t)i * 3); sp[i] = __ldg(pp + (int64_t)i * 3); }
  const float k10 = 10.f / logf(10.f);
  float acc = 0.f;
  for (int f = 0; f < n_frames; ++f) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int j = f * mel.hop + i - n / 2;                  // centred frame over the reflect-padded signal (np.pad mode='reflect')
      if (j < 0) j = -j;
      if (j >= t) j = 2 * (t - 1) - j;
      const float w = __ldg(P.hann + i);
      b0[i] = make_float2(sg[j] * w, 0.f);
      b0[n + i] = make_float2(sp[j] * w, 0.f);
    }
    const float2* r = block_fft_nf(b0, b1, P, 2);
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
      pw[k] = r[k].x * r[k].x + r[k].y * r[k].y;
      pw[nb + k] = r[n + k].x * r[n + k].x + r[n + k].y * r[n + k].y;
    }
    __syncthreads();
    if (threadIdx.x < mel.n_mels) {
      const int s0 = __ldg(mel.start + threadIdx.x), ln = __ldg(mel.len + threadIdx.x);
      const float* w = mel.w + __ldg(mel.off + threadIdx.x);
      float mg = 0.f, mp = 0.f;
      for (int k = 0; k < ln; ++k) {
        const float wk = __ldg(w + k);
        mg = fmaf(wk, pw[s0 + k], mg);
        mp = fmaf(wk, pw[nb + s0 + k], mp);
      }
      const float d = k10 * (logf(fabsf(mg) + 1e-2f) - logf(fabsf(mp) + 1e-2f));      // myutils.py:98-100
      acc = fmaf(d, d, acc);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[wc] = sqrtf(acc / (float)(mel.n_mels * n_frames));           // myutils.py:105
}

int launch_mel_lsd(const float* pred, const float* gt, int batch, int t, int audio_rate, float* out, cudaStream_t st) {
  SAG_REQUIRE(batch > 0 && audio_rate > 0, SAG_EINVAL, "mel_lsd: bad arguments");
  MelBank mel;
  SAG_TRY(get_mel_bank(audio_rate, &mel));
  SAG_REQUIRE(t > mel.n_fft / 2, SAG_EINVAL, "mel_lsd: %d samples are too few for reflect padding by %d", t, mel.n_fft / 2);
  FftPlan P;
  SAG_TRY(get_plan(mel.n_fft, &P));
  const int n_frames = 1 + t / mel.hop;                 // librosa 0.6 stft with center=True
  const size_t smem = (size_t)4 * mel.n_fft * sizeof(float2) + (size_t)2 * t * sizeof(float) + (size_t)2 * (mel.n_fft / 2 + 1) * sizeof(float);
  SAG_REQUIRE(smem <= 220 * 1024, SAG_EUNSUPPORTED, "mel_lsd: %zu bytes of shared memory needed", smem);
  SAG_CHECK_CUDA(cudaFuncSetAttribute(mel_lsd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mel_lsd_kernel<<<batch * 3, 256, smem, st>>>(pred, gt, t, P, mel, n_frames, out);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}


// ---- spherical-harmonic projection + RMS map -------------------------------------------------------------------
struct ShMesh {
  int n_nu, n_phi;
  const float4* Y;    // [n_nu*n_phi] (W,Y,Z,X) coefficients, ACN/SN3D order 1
};
static std::mutex g_mesh_mu;
static std::map<std::pair<int, int>, ShMesh> g_meshes;   // (device, ang_res*1000)

static void mesh_dims(double ang_res, int* n_nu, int* n_phi) {
  // np.arange(start, stop, step) has ceil((stop-start)/step) elements (distance.py:10-11)
  *n_phi = (int)ceil((180.0 - (-180.0)) / ang_res);
  *n_nu = (int)ceil((90.1 - (-90.0)) / ang_res);
}

static int get_mesh(float ang_res_f, ShMesh* out) {
  int dev = 0;
  SAG_CHECK_CUDA(cudaGetDevice(&dev));
  SAG_REQUIRE(ang_res_f > 0.f, SAG_EINVAL, "sh_rms: angular resolution must be positive");
  std::lock_guard<std::mutex> lk(g_mesh_mu);
  auto key = std::make_pair(dev, (int)lround((double)ang_res_f * 1000.0));
  auto it = g_meshes.find(key);
  if (it != g_meshes.end()) { *out = it->second; return SAG_OK; }
  const double res = (double)ang_res_f;
  ShMesh m;
  mesh_dims(res, &m.n_nu, &m.n_phi);
  std::vector<float4> Y((size_t)m.n_nu * m.n_phi);
  for (int iv = 0; iv < m.n_nu; ++iv)
    for (int ip = 0; ip < m.n_phi; ++ip) {
      // distance.py:10-11: phi = flip(arange(-180,180,res))/180*pi ; nu = arange(-90,90.1,res)/180*pi
      double phi = (-180.0 + res * (double)(m.n_phi - 1 - ip)) / 180.0 * M_PI;
      double nu = (-90.0 + res * (double)iv) / 180.0 * M_PI;
      // position.py:23-37: polar -> cartesian -> polar round trip
      double x = cos(phi) * cos(nu), y = sin(phi) * cos(nu), z = sin(nu);
      double phi2 = atan2(y, x), nu2 = atan2(z, sqrt(x * x + y * y));
      // common.py:151-157 at order 1, ACN/SN3D: [1, sin(phi)cos(nu), sin(nu), cos(phi)cos(nu)]
      double cn = sqrt(fmax(0.0, 1.0 - sin(nu2) * sin(nu2)));     // -lpmv(1,1,sin nu) = sqrt(1-sin^2 nu)
      Y[(size_t)iv * m.n_phi + ip] = make_float4(1.f, (float)(sin(phi2) * cn), (float)sin(nu2), (float)(cos(phi2) * cn));
    }
  float4* d = nullptr;
  SAG_CHECK_CUDA(cudaMalloc(&d, sizeof(float4) * Y.size()));
  SAG_CHECK_CUDA(cudaMemcpy(d, Y.data(), sizeof(float4) * Y.size(), cudaMemcpyHostToDevice));
  m.Y = d;
  g_meshes[key] = m;
  *out = m;
  return SAG_OK;
}

int sh_mesh_dims(float ang_res, int* n_nu, int* n_phi) {
  if (!(ang_res > 0.f)) { set_error("sh_rms: angular resolution must be positive"); return SAG_EINVAL; }
  mesh_dims((double)ang_res, n_nu, n_phi);
  return SAG_OK;
}

// grid (ceil(D/dirs_per_block), batch); each warp owns one direction at a time and strides over time.
__global__ void __launch_bounds__(256) sh_rms_kernel(const float4* __restrict__ ambi, int t, const ShMesh mesh,
                                                     float* __restrict__ rms) {
  const int b = blockIdx.y;
  const int D = mesh.n_nu * mesh.n_phi;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float4* a = ambi + (int64_t)b * t;
  for (int d = blockIdx.x * nw + wid; d < D; d += gridDim.x * nw) {
    const float4 y = __ldg(mesh.Y + d);
    float acc = 0.f;
    for (int i = lane; i < t; i += 32) {
      float4 v = __ldg(a + i);
      float s = v.x * y.x + v.y * y.y + v.z * y.z + v.w * y.w;     // decoder.py:26
      acc = fmaf(s, s, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      int iv = d / mesh.n_phi, ip = d % mesh.n_phi;
      rms[((int64_t)b * mesh.n_nu + (mesh.n_nu - 1 - iv)) * mesh.n_phi + ip] = sqrtf(acc / (float)t);   // flipud
    }
  }
}

int launch_sh_rms(const float* ambi, int batch, int t, float ang_res, float* rms, cudaStream_t st) {
  SAG_REQUIRE(batch > 0 && t > 0, SAG_EINVAL, "sh_rms: bad arguments");
  ShMesh m;
  SAG_TRY(get_mesh(ang_res, &m));
  const int D = m.n_nu * m.n_phi;
  dim3 grid(std::min(cdiv(D, 8), 64), batch);
  sh_rms_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(ambi), t, m, rms);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

}  // namespace sag
