// fp32 gather-GEMM on the FFMA pipe: the SAG_PREC_FP32 (parity) implementation of every dense contraction
// on the path -- tfw.conv_2d (reference core.py:156-220), the output phases of tfw.deconv_2d
// (core.py:96-153) and tfw.fully_connected (core.py:43-93) all reduce to the GatherGeom of common.cuh.
//
//   C[m, co] = sum_{t, ci} X[pixel(m) + tap t, ci] * Wt[widx[t]][ci][co]       (zero outside the image)
//
// Tile 128 x BN x 16, 256 threads, 8 x TN register micro-tile, register prefetch of the next K chunk.
// Roofline: FFMA-bound (2*128*BN*16 flop per 16*(128+BN)*4 B of smem traffic); this is the exact-fp32
// reference path, the tensor-core path is conv_umma.cu.
#include "common.cuh"

namespace sag {

constexpr int BM = 128, BK = 16, TM = 8, NTHREADS = 256;

template <int BN, int TN, bool VECA>
__global__ void __launch_bounds__(NTHREADS) gather_gemm_ffma_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ w,
                                                                    float* __restrict__ y,
                                                                    const __grid_constant__ GatherGeom g,
                                                                    const Epilogue ep) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ float s_sum[BN], s_sqs[BN];

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  const int K = g.T * g.Cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;

  // ---- A loader role: row ar = tid % 128, k half ah = tid / 128 (8 consecutive k) ----
  const int ar = tid & (BM - 1);
  const int ah = tid >> 7;
  const int64_t am = m0 + ar;
  const bool arow_ok = am < M;
  int a_iy0 = 0, a_ix0 = 0;
  const float* a_img = x;
  if (arow_ok) {
    int j = (int)(am % g.PW);
    int64_t r = am / g.PW;
    int i = (int)(r % g.PH);
    int n = (int)(r / g.PH);
    a_iy0 = i * g.isy;
    a_ix0 = j * g.isx;
    a_img = x + (int64_t)n * g.H * g.W * g.x_ld;
  }
  // ---- B loader role: k row bk = tid / (BN/4), 4 columns at bc = (tid % (BN/4))*4 ----
  constexpr int BQ = BN / 4;
  const int bk = tid / BQ;
  const int bc = (tid % BQ) * 4;
  const bool b_active = bk < BK;
  const bool b_vec = (g.Cout % 4) == 0;

  float areg[8];
  float breg[4];

  auto load_chunk = [&](int k0) {
    // A
    int kk = k0 + ah * 8;
    if (VECA) {
      // Cin % 8 == 0: the 8 consecutive k share one tap and are contiguous in memory
      bool ok = arow_ok && kk < K;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (ok) {
        int t = kk / g.Cin;
        int ci = kk - t * g.Cin;
        int iy = a_iy0 + g.dy[t], ix = a_ix0 + g.dx[t];
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
          const float4* p = reinterpret_cast<const float4*>(a_img + ((int64_t)iy * g.W + ix) * g.x_ld + ci);
          v0 = __ldg(p);
          v1 = __ldg(p + 1);
        }
      }
      areg[0] = v0.x; areg[1] = v0.y; areg[2] = v0.z; areg[3] = v0.w;
      areg[4] = v1.x; areg[5] = v1.y; areg[6] = v1.z; areg[7] = v1.w;
    } else {
      int t = kk / g.Cin;
      int ci = kk - t * g.Cin;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = 0.f;
        if (arow_ok && (kk + e) < K) {
          int iy = a_iy0 + g.dy[t], ix = a_ix0 + g.dx[t];
          if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = __ldg(a_img + ((int64_t)iy * g.W + ix) * g.x_ld + ci);
        }
        areg[e] = v;
        if (++ci == g.Cin) { ci = 0; ++t; }
      }
    }
    // B
    if (b_active) {
      int k = k0 + bk;
      breg[0] = breg[1] = breg[2] = breg[3] = 0.f;
      if (k < K) {
        int t = k / g.Cin;
        int ci = k - t * g.Cin;
        const float* p = w + ((int64_t)g.widx[t] * g.Cin + ci) * g.Cout + co0 + bc;
        if (b_vec && co0 + bc + 3 < g.Cout) {
          float4 v = __ldg(reinterpret_cast<const float4*>(p));
          breg[0] = v.x; breg[1] = v.y; breg[2] = v.z; breg[3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (co0 + bc + e < g.Cout) breg[e] = __ldg(p + e);
        }
      }
    }
  };
  auto store_chunk = [&]() {
#pragma unroll
    for (int e = 0; e < 8; ++e) As[ah * 8 + e][ar] = areg[e];
    if (b_active) *reinterpret_cast<float4*>(&Bs[bk][bc]) = make_float4(breg[0], breg[1], breg[2], breg[3]);
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[TM][TN];
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;

  const int nchunks = (K + BK - 1) / BK;
  load_chunk(0);
  store_chunk();
  __syncthreads();
  for (int kc = 0; kc < nchunks; ++kc) {
    if (kc + 1 < nchunks) load_chunk((kc + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int c = 0; c < TN; ++c) b[c] = Bs[k][tx * TN + c];
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
    if (kc + 1 < nchunks) {
      store_chunk();
      __syncthreads();
    }
  }

  // ---- epilogue: bias, activation, strided store, optional batch-norm statistics ----
  const bool do_stats = ep.stat_sum != nullptr;
  if (do_stats) {
    for (int c = tid; c < BN; c += NTHREADS) { s_sum[c] = 0.f; s_sqs[c] = 0.f; }
    __syncthreads();
  }
  float bsv[TN];
#pragma unroll
  for (int c = 0; c < TN; ++c) {
    int co = co0 + tx * TN + c;
    bsv[c] = (ep.bias != nullptr && co < g.Cout) ? __ldg(ep.bias + co) : 0.f;
  }
  float psum[TN], psqs[TN];
#pragma unroll
  for (int c = 0; c < TN; ++c) { psum[c] = 0.f; psqs[c] = 0.f; }
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    int64_t m = m0 + ty * TM + r;
    if (m >= M) continue;
    int j = (int)(m % g.PW);
    int64_t q = m / g.PW;
    int i = (int)(q % g.PH);
    int n = (int)(q / g.PH);
    float* yp = y + (int64_t)n * g.y_sn + (int64_t)(g.oy0 + i * g.osy) * g.y_sh + (int64_t)(g.ox0 + j * g.osx) * g.y_sw;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      int co = co0 + tx * TN + c;
      if (co < g.Cout) {
        float v = acc[r][c] + bsv[c];
        if (ep.relu) v = fmaxf(v, 0.f);
        yp[(int64_t)co * g.y_sc] = v;
        psum[c] += v;
        psqs[c] += v * v;
      }
    }
  }
  if (do_stats) {
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      atomicAdd(&s_sum[tx * TN + c], psum[c]);
      atomicAdd(&s_sqs[tx * TN + c], psqs[c]);
    }
    __syncthreads();
    for (int c = tid; c < BN; c += NTHREADS) {
      if (co0 + c < g.Cout) {
        fx_atomic_add(ep.stat_sum + 2 * (co0 + c), s_sum[c]);
        fx_atomic_add(ep.stat_sqs + 2 * (co0 + c), s_sqs[c]);
      }
    }
  }
}

int launch_gather_gemm_ffma(const float* x, const float* w, float* y, const GatherGeom& g, const Epilogue& ep,
                            cudaStream_t st) {
  SAG_REQUIRE(g.T >= 0 && g.T <= kMaxTaps, SAG_EINVAL, "gather_gemm: %d taps unsupported (max %d)", g.T, kMaxTaps);
  int64_t M = (int64_t)g.N * g.PH * g.PW;
  if (M == 0) return SAG_OK;
  bool veca = (g.Cin % 8 == 0) && (g.x_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  SAG_REQUIRE(cdiv64(M, BM) < (1ll << 31), SAG_EINVAL, "gather_gemm: too many rows");
  if (g.Cout <= 32) {
    dim3 grid((unsigned)cdiv64(M, BM), cdiv(g.Cout, 32));
    if (veca) gather_gemm_ffma_kernel<32, 2, true><<<grid, NTHREADS, 0, st>>>(x, w, y, g, ep);
    else gather_gemm_ffma_kernel<32, 2, false><<<grid, NTHREADS, 0, st>>>(x, w, y, g, ep);
  } else {
    dim3 grid((unsigned)cdiv64(M, BM), cdiv(g.Cout, 64));
    if (veca) gather_gemm_ffma_kernel<64, 4, true><<<grid, NTHREADS, 0, st>>>(x, w, y, g, ep);
    else gather_gemm_ffma_kernel<64, 4, false><<<grid, NTHREADS, 0, st>>>(x, w, y, g, ep);
  }
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

}  // namespace sag
