// tcgen05 gather-GEMM: the tensor-core implementation of every dense contraction on the path -- tfw.conv_2d
// (reference core.py:156-220), tfw.deconv_2d (core.py:96-153, as one sub-pixel GEMM per layer) and
// tfw.fully_connected (core.py:43-93).  im2col-free: the A operand is gathered straight from the NHWC activation.
//
//   C[m, n] = sum_{t, ci} X[pixel(m) + tap t, ci] * Wk[t*Cin + ci, n]          (zero outside the image)
//
// CTA = one 128 x BN output tile, 9 warps:
//   warps 0-7  producers: gather fp32 activations (coalesced 32 B per thread), split each value into bf16 hi + bf16 lo,
//              store both planes into shared memory in the UMMA K-major SWIZZLE_128B canonical layout; warp 0 lane 0 also
//              issues the bulk-async (TMA, cp.async.bulk) copy of the pre-packed weight tile; afterwards the same warps run
//              the epilogue: tcgen05.ld of the accumulator, bias / ReLU, batch-norm statistics, strided or mapped store.
//   warp 8     allocates TMEM and (one elected lane) issues tcgen05.mma kind::f16 with the accumulator in TMEM:
//              SAG_PREC_BF16   : 1 MMA per K step  (A_hi x B_hi)
//              SAG_PREC_BF16X3 : 3 MMAs per K step (A_hi x B_hi + A_lo x B_hi + A_hi x B_lo) -> fp32-grade products
// full/empty mbarrier ring between producers and the MMA issuer, tcgen05.commit releases stages and publishes the
// accumulator.  Roofline: tensor pipe (BN=128, x3: 768 clk of MMA per 64-wide K chunk against ~650 clk to gather the
// 32 KB A chunk from L2); the A gather re-reads each activation once per tap from L2, never from HBM.
#include "model.cuh"
#include <cuda_bf16.h>

namespace sag {

namespace {

constexpr int UM_BM = 128;                  // rows per tile (UMMA M, one TMEM lane per row)
constexpr int UM_BK = 64;                   // K elements per stage = one 128-byte swizzle atom of bf16
constexpr int UM_PRODUCER_WARPS = 8;
constexpr int UM_PRODUCERS = UM_PRODUCER_WARPS * 32;
constexpr int UM_THREADS = UM_PRODUCERS + 32;
constexpr int UM_A_PLANE = UM_BM * 128;     // bytes of one A plane per stage
constexpr int UM_MAX_STAGES = 8;
constexpr int UM_BAR_BYTES = 256;

struct UmmaArgs {
  const float* x;
  const uint8_t* wpacked;
  float* y;
  const float* bias;      // indexed by GEMM column (already expanded for mapped outputs) or null
  int relu;
  double* stat_sum;
  double* stat_sqs;
  const int* col_off;     // mapped output (sub-pixel transposed conv): element offset per column, or null
  const short* col_dy;
  const short* col_dx;
  int oh_lim, ow_lim;     // validity window of mapped outputs
  int Ntot;               // valid GEMM columns
  int K;                  // valid K
  int KC;                 // K chunks of 64
  int stages;
  int vec_store;          // groups of 4 columns are contiguous and 16-byte aligned in the output
  float* partial;         // split-K: raw accumulators [split][M][n_pad] (bias / activation / statistics run in the reduce)
  int n_pad;
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16 operands, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO),
// descriptor version 1 (sm_100), layout type 2.  `addr` may be advanced by 32 bytes per UMMA_K inside the atom.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// fp32 -> bf16 hi (round to nearest even) and the bf16 of the remainder; 8 values -> two 16-byte vectors
__device__ __forceinline__ void split8(const float* f, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    float r0 = f[2 * i] - __low2float(hh), r1 = f[2 * i + 1] - __high2float(hh);
    __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    l[i] = *reinterpret_cast<uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// sum of v[c] over the 32 lanes for 16 columns at once (transposing butterfly, 16 shuffles): afterwards the lanes
// with an even index hold the total of column ((lane>>4)&1)*8 + ((lane>>3)&1)*4 + ((lane>>2)&1)*2 + ((lane>>1)&1).
__device__ __forceinline__ float warp_colsum16(const float* v, int lane) {
  float a[8];
  const bool u16 = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float send = u16 ? v[i] : v[i + 8];
    float keep = u16 ? v[i + 8] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float b[4];
  const bool u8 = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = u8 ? a[i] : a[i + 4];
    float keep = u8 ? a[i + 4] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float c[2];
  const bool u4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = u4 ? b[i] : b[i + 2];
    float keep = u4 ? b[i + 2] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const bool u2 = lane & 2;
  float send = u2 ? c[0] : c[1];
  float keep = u2 ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// VEC: every 8-element K group lies inside one tap and is contiguous + 16-byte aligned in memory (Cin % 8 == 0).
// PF: K chunks whose global loads are in flight per producer thread (register prefetch distance).
template <int BN, int NSPLIT, bool VEC, int PF>
__global__ void __launch_bounds__(UM_THREADS, (BN <= 64 && PF <= 2 ? 2 : 1))
gather_gemm_umma_kernel(const __grid_constant__ GatherGeom g, const UmmaArgs a) {
  constexpr int PLANES = NSPLIT == 3 ? 2 : 1;
  constexpr int B_PLANE = BN * 128;
  constexpr int STAGE_BYTES = PLANES * (UM_A_PLANE + B_PLANE);
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);

  extern __shared__ __align__(16) uint8_t um_smem[];
  __shared__ float s_sum[BN], s_sqs[BN];
  __shared__ int s_any_valid;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bars = smem_u32(um_smem);
  const uint32_t tiles = (bars + UM_BAR_BYTES + 1023u) & ~1023u;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * UM_MAX_STAGES, bar_acc = bars + 16 * UM_MAX_STAGES;
  const uint32_t tmem_slot = bar_acc + 8;
  const int S = a.stages;
  // split-K: this CTA owns K chunks [kc_begin, kc_end)
  const int kc_begin = (int)((int64_t)blockIdx.z * a.KC / gridDim.z);
  const int kc_end = (int)((int64_t)(blockIdx.z + 1) * a.KC / gridDim.z);
  const int KC = kc_end - kc_begin;

  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  const int64_t m0 = (int64_t)blockIdx.x * UM_BM;
  const int nt = blockIdx.y;
  const int n_base = nt * BN;

  // ---- mapped outputs: skip tiles none of whose (row, column) pairs land inside the output window ----
  if (a.col_off != nullptr) {
    if (tid == 0) s_any_valid = 0;
    __syncthreads();
    if (tid < UM_BM && m0 + tid < M) {
      int64_t m = m0 + tid;
      int j = (int)(m % g.PW);
      int i = (int)((m / g.PW) % g.PH);
      const int oy = g.oy0 + i * g.osy, ox = g.ox0 + j * g.osx;
      bool any = false;
      for (int c = 0; c < BN && !any; ++c) {
        int n = n_base + c;
        if (n >= a.Ntot) break;
        any = (unsigned)(oy + a.col_dy[n]) < (unsigned)a.oh_lim && (unsigned)(ox + a.col_dx[n]) < (unsigned)a.ow_lim;
      }
      if (any) s_any_valid = 1;
    }
    __syncthreads();
    if (!s_any_valid) return;
  }

  // ---- one-time setup ----
  if (tid < BN) { s_sum[tid] = 0.f; s_sqs[tid] = 0.f; }
  if (warp == UM_PRODUCER_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(bar_full + 8 * s, UM_PRODUCER_WARPS + 1);   // 8 producer warps + the expect_tx arrival of the B copy
        mbar_init(bar_empty + 8 * s, 1);                       // one tcgen05.commit
      }
      mbar_init(bar_acc, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  if (warp < UM_PRODUCER_WARPS) {
    // ================================ producers ================================
    const int jchunk = tid & 7;                 // which 8-element (16-byte bf16) group of the 64-wide K chunk
    int iy0[4], ix0[4];
    const float* img[4];
    bool rok[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 32 + (tid >> 3);
      const int64_t m = m0 + r;
      rok[it] = m < M;
      iy0[it] = 0; ix0[it] = 0; img[it] = a.x;
      if (rok[it]) {
        int j = (int)(m % g.PW);
        int64_t q = m / g.PW;
        int i = (int)(q % g.PH);
        int n = (int)(q / g.PH);
        iy0[it] = i * g.isy;
        ix0[it] = j * g.isx;
        img[it] = a.x + (int64_t)n * g.H * g.W * g.x_ld;
      }
    }
    // loads of one K chunk into registers (zero outside the image / beyond K)
    auto load_chunk = [&](int kc, float (&f)[4][8]) {
      if (kc >= kc_end) return;
      const int kk = kc * UM_BK + jchunk * 8;
      if (VEC) {
        const int t = kk / g.Cin;
        const int ci = kk - t * g.Cin;
        const bool kok = kk < a.K;
        const int dy = kok ? g.dy[t] : 0, dx = kok ? g.dx[t] : 0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int iy = iy0[it] + dy, ix = ix0[it] + dx;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (kok && rok[it] && (unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) {
            const float4* p = reinterpret_cast<const float4*>(img[it] + ((int64_t)iy * g.W + ix) * g.x_ld + ci);
            v0 = __ldg(p);
            v1 = __ldg(p + 1);
          }
          f[it][0] = v0.x; f[it][1] = v0.y; f[it][2] = v0.z; f[it][3] = v0.w;
          f[it][4] = v1.x; f[it][5] = v1.y; f[it][6] = v1.z; f[it][7] = v1.w;
        }
      } else {
        const int t = kk / g.Cin;
        const int ci0 = kk - t * g.Cin;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          int tt = t, ci = ci0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = 0.f;
            if (rok[it] && kk + e < a.K) {
              const int iy = iy0[it] + g.dy[tt], ix = ix0[it] + g.dx[tt];
              if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W)
                v = __ldg(img[it] + ((int64_t)iy * g.W + ix) * g.x_ld + ci);
            }
            f[it][e] = v;
            if (++ci == g.Cin) { ci = 0; ++tt; }
          }
        }
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    // split hi/lo, store into the swizzled stage, publish it
    auto store_chunk = [&](int kc, float (&f)[4][8]) {
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);     // the MMAs that read this stage last time round have completed
      const uint32_t st_base = tiles + (uint32_t)stage * STAGE_BYTES;
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_full + 8 * stage, PLANES * B_PLANE);
        bulk_g2s(st_base + PLANES * UM_A_PLANE, a.wpacked + ((size_t)nt * a.KC + kc) * (size_t)(PLANES * B_PLANE),
                 PLANES * B_PLANE, bar_full + 8 * stage);
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 32 + (tid >> 3);
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((jchunk ^ (r & 7)) << 4);
        uint4 hi, lo;
        split8(f[it], hi, lo);
        st_shared_v4(st_base + off, hi);
        if (PLANES == 2) st_shared_v4(st_base + UM_A_PLANE + off, lo);
      }
      fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
      if (++stage == S) { stage = 0; phase ^= 1; }
    };
    float f[PF][4][8];
#pragma unroll
    for (int p = 0; p < PF; ++p) load_chunk(kc_begin + p, f[p]);
#pragma unroll 1
    for (int kc = kc_begin; kc < kc_end; kc += PF) {
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        if (kc + p < kc_end) {
          store_chunk(kc + p, f[p]);
          load_chunk(kc + p + PF, f[p]);
        }
      }
    }

    // ================================ epilogue ================================
    if (KC > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const int64_t m = m0 + r;
    const bool row_ok = m < M;
    float* yrow = a.y;
    int oy = 0, ox = 0;
    if (row_ok) {
      int j = (int)(m % g.PW);
      int64_t qq = m / g.PW;
      int i = (int)(qq % g.PH);
      int n = (int)(qq / g.PH);
      oy = g.oy0 + i * g.osy;
      ox = g.ox0 + j * g.osx;
      yrow = a.y + (int64_t)n * g.y_sn + (int64_t)oy * g.y_sh + (int64_t)ox * g.y_sw;
    }
    const bool do_stats = a.stat_sum != nullptr && a.partial == nullptr;
    constexpr int HALF = BN / 2;
#pragma unroll 1
    for (int c0 = half * HALF; c0 < (half + 1) * HALF; c0 += 16) {
      float v[16];
      if (KC > 0) {
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0.f;
      }
      const int n0 = n_base + c0;
      if (a.partial != nullptr) {     // split-K: raw accumulators, finished by splitk_reduce_kernel
        if (row_ok) {
          float4* pp = reinterpret_cast<float4*>(a.partial + ((int64_t)blockIdx.z * M + m) * a.n_pad + n0);
#pragma unroll
          for (int e = 0; e < 4; ++e) pp[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
        continue;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int n = n0 + e;
        float b = (a.bias != nullptr && n < a.Ntot) ? __ldg(a.bias + n) : 0.f;
        float w = v[e] + b;
        if (a.relu) w = fmaxf(w, 0.f);
        v[e] = w;
      }
      if (row_ok) {
        if (a.col_off == nullptr) {
          if (a.vec_store) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              if (n0 + e < a.Ntot) *reinterpret_cast<float4*>(yrow + (n0 + e)) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (n0 + e < a.Ntot) yrow[(int64_t)(n0 + e) * g.y_sc] = v[e];
          }
        } else {
          if (a.vec_store) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const int n = n0 + e;
              if (n < a.Ntot && (unsigned)(oy + a.col_dy[n]) < (unsigned)a.oh_lim &&
                  (unsigned)(ox + a.col_dx[n]) < (unsigned)a.ow_lim)
                *reinterpret_cast<float4*>(yrow + a.col_off[n]) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int n = n0 + e;
              if (n < a.Ntot && (unsigned)(oy + a.col_dy[n]) < (unsigned)a.oh_lim &&
                  (unsigned)(ox + a.col_dx[n]) < (unsigned)a.ow_lim)
                yrow[a.col_off[n]] = v[e];
            }
          }
        }
      }
      if (do_stats) {   // batch-norm statistics of the stored values (padded rows / columns contribute zeros)
        float sq[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (!row_ok || n0 + e >= a.Ntot) v[e] = 0.f;
          sq[e] = v[e] * v[e];
        }
        const float cs = warp_colsum16(v, lane);
        const float cq = warp_colsum16(sq, lane);
        if ((lane & 1) == 0) {
          const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          atomicAdd(&s_sum[c0 + col], cs);
          atomicAdd(&s_sqs[c0 + col], cq);
        }
      }
    }
    tc_fence_before();
  } else {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t st_base = tiles + (uint32_t)stage * STAGE_BYTES;
        const uint32_t a_hi = st_base, a_lo = st_base + UM_A_PLANE;
        const uint32_t b_hi = st_base + PLANES * UM_A_PLANE, b_lo = b_hi + B_PLANE;
#pragma unroll
        for (int k4 = 0; k4 < UM_BK / 16; ++k4) {
          const uint64_t da_hi = make_sw128_desc(a_hi + k4 * 32), db_hi = make_sw128_desc(b_hi + k4 * 32);
          umma_bf16(tmem_acc, da_hi, db_hi, IDESC, (kc > kc_begin || k4 > 0) ? 1u : 0u);
          if (NSPLIT == 3) {
            const uint64_t da_lo = make_sw128_desc(a_lo + k4 * 32), db_lo = make_sw128_desc(b_lo + k4 * 32);
            umma_bf16(tmem_acc, da_lo, db_hi, IDESC, 1u);
            umma_bf16(tmem_acc, da_hi, db_lo, IDESC, 1u);
          }
        }
        umma_commit(bar_empty + 8 * stage);      // frees the stage once these MMAs have read it
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (KC > 0) umma_commit(bar_acc);          // accumulator complete
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == UM_PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, TMEM_COLS);
  }
  if (a.stat_sum != nullptr && a.partial == nullptr && tid < BN && n_base + tid < a.Ntot) {
    atomicAdd(a.stat_sum + n_base + tid, (double)s_sum[tid]);
    atomicAdd(a.stat_sqs + n_base + tid, (double)s_sqs[tid]);
  }
}

// ---- weight packing: Wk fp32 [K][N] (row stride ldw) -> per (N tile, K chunk) bf16 hi (+lo) planes in the swizzled
//      K-major layout the MMA reads, so a stage's B operand is one contiguous bulk copy ----
__global__ void umma_pack_weights_kernel(const float* __restrict__ wk, int K, int N, int64_t ldw, int BN, int KC, int NT,
                                         int planes, uint8_t* __restrict__ out) {
  const int64_t total = (int64_t)NT * KC * BN * 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int nl = (int)(idx % BN);
    int64_t r = idx / BN;
    const int j = (int)(r % 8);
    r /= 8;
    const int kc = (int)(r % KC);
    const int nt = (int)(r / KC);
    const int n = nt * BN + nl;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * UM_BK + j * 8 + e;
      f[e] = (k < K && n < N) ? __ldg(wk + (int64_t)k * ldw + n) : 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * i] - __low2float(hh), f[2 * i + 1] - __high2float(hh));
      h[i] = *reinterpret_cast<uint32_t*>(&hh);
      l[i] = *reinterpret_cast<uint32_t*>(&ll);
    }
    const size_t plane_bytes = (size_t)BN * 128;
    uint8_t* tile = out + ((size_t)nt * KC + kc) * (size_t)planes * plane_bytes;
    const size_t off = (size_t)nl * 128 + (size_t)((j ^ (nl & 7)) << 4);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(tile + plane_bytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// tf.nn.conv2d_transpose weights [kh,kw,Cout,Cin] (core.py:118) -> sub-pixel GEMM matrix Wk[(a*tx+b)*Cin + ci][n]:
//   n = (py*sw + px)*Cout + co   (order 0, NHWC outputs)   or   n = (py*Cout + co)*sw + px   (order 1, planar outputs)
//   value = w[py + sh*a, px + sw*b, co, ci], zero where the kernel index falls outside [0,kh) x [0,kw).
__global__ void subpixel_weights_kernel(const float* __restrict__ w, int kh, int kw, int cout, int cin, int sh, int sw,
                                        int ty, int tx, int order, float* __restrict__ wk) {
  const int N = sh * sw * cout;
  const int64_t total = (int64_t)ty * tx * cin * N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int n = (int)(idx % N);
    int64_t r = idx / N;
    const int ci = (int)(r % cin);
    const int t = (int)(r / cin);
    const int ta = t / tx, tb = t % tx;
    int py, px, co;
    if (order == 0) { co = n % cout; int ph = n / cout; px = ph % sw; py = ph / sw; }
    else { px = n % sw; int q = n / sw; co = q % cout; py = q / cout; }
    const int p = py + sh * ta, qq = px + sw * tb;
    wk[idx] = (p < kh && qq < kw) ? __ldg(w + (((int64_t)p * kw + qq) * cout + co) * cin + ci) : 0.f;
  }
}

// HWIO conv weights [kh][kw][cin][cout] -> [kh][kw2][cin2][cout] with zeros in the added taps / channels
__global__ void expand_hwio_kernel(const float* __restrict__ w, int kh, int kw, int cin, int cout, int kw2, int cin2,
                                   float* __restrict__ out) {
  const int64_t total = (int64_t)kh * kw2 * cin2 * cout;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int co = (int)(idx % cout);
    int64_t r = idx / cout;
    const int c = (int)(r % cin2); r /= cin2;
    const int sx = (int)(r % kw2);
    const int ry = (int)(r / kw2);
    out[idx] = (c < cin && sx < kw) ? __ldg(w + (((int64_t)ry * kw + sx) * cin + c) * cout + co) : 0.f;
  }
}

__global__ void expand_bias_kernel(const float* __restrict__ bias, int cout, int sw, int N, int order, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int co = order == 0 ? n % cout : (n / sw) % cout;
  out[n] = __ldg(bias + co);
}

// ---- split-K finish: sum the partial accumulators, then the same bias / activation / statistics / store as the
//      fused epilogue.  One thread per (row, 4-column group); a block owns `rows_per_block` consecutive rows. ----
__global__ void __launch_bounds__(128) splitk_reduce_kernel(const __grid_constant__ GatherGeom g, const UmmaArgs a, int Z,
                                                            int rows_per_block) {
  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  const int ncg = (a.Ntot + 3) / 4;
  for (int cg = threadIdx.x; cg < ncg; cg += blockDim.x) {
    const int n = cg * 4;
    float bs[4], ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssqs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) bs[e] = (a.bias != nullptr && n + e < a.Ntot) ? __ldg(a.bias + n + e) : 0.f;
    for (int64_t m = r0; m < r1; ++m) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int z = 0; z < Z; ++z) {
        const float4 p = __ldcs(reinterpret_cast<const float4*>(a.partial + ((int64_t)z * M + m) * a.n_pad + n));
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
      }
      float v[4] = {acc.x + bs[0], acc.y + bs[1], acc.z + bs[2], acc.w + bs[3]};
      if (a.relu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
      }
      const int j = (int)(m % g.PW);
      const int64_t q = m / g.PW;
      const int i = (int)(q % g.PH);
      const int b = (int)(q / g.PH);
      const int oy = g.oy0 + i * g.osy, ox = g.ox0 + j * g.osx;
      float* yrow = a.y + (int64_t)b * g.y_sn + (int64_t)oy * g.y_sh + (int64_t)ox * g.y_sw;
      if (a.col_off == nullptr) {
        if (a.vec_store) {
          *reinterpret_cast<float4*>(yrow + n) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < a.Ntot) yrow[(int64_t)(n + e) * g.y_sc] = v[e];
        }
      } else {
        if (a.vec_store) {
          if ((unsigned)(oy + a.col_dy[n]) < (unsigned)a.oh_lim && (unsigned)(ox + a.col_dx[n]) < (unsigned)a.ow_lim)
            *reinterpret_cast<float4*>(yrow + a.col_off[n]) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < a.Ntot && (unsigned)(oy + a.col_dy[n + e]) < (unsigned)a.oh_lim &&
                (unsigned)(ox + a.col_dx[n + e]) < (unsigned)a.ow_lim)
              yrow[a.col_off[n + e]] = v[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { ssum[e] += v[e]; ssqs[e] = fmaf(v[e], v[e], ssqs[e]); }
    }
    if (a.stat_sum != nullptr && r1 > r0) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < a.Ntot) {
          atomicAdd(a.stat_sum + n + e, (double)ssum[e]);
          atomicAdd(a.stat_sqs + n + e, (double)ssqs[e]);
        }
    }
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int BN, int NSPLIT, bool VEC, int PF>
int launch_cfg(const GatherGeom& g, const UmmaArgs& a_in, int nt, int Z, cudaStream_t st) {
  constexpr int PLANES = NSPLIT == 3 ? 2 : 1;
  constexpr int STAGE_BYTES = PLANES * (UM_A_PLANE + BN * 128);
  constexpr bool TWO_PER_SM = BN <= 64 && PF <= 2;
  UmmaArgs a = a_in;
  const int budget = (TWO_PER_SM ? 110 : 220) * 1024 - UM_BAR_BYTES - 1024;
  int S = budget / STAGE_BYTES;
  if (S > 4) S = 4;
  if (S < 2) S = 2;
  a.stages = S;
  const size_t smem = (size_t)UM_BAR_BYTES + 1024 + (size_t)S * STAGE_BYTES;
  auto kern = gather_gemm_umma_kernel<BN, NSPLIT, VEC, PF>;
  static bool attr_set = false;
  if (!attr_set) {
    SAG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  dim3 grid((unsigned)cdiv64(M, UM_BM), (unsigned)nt, (unsigned)Z);
  kern<<<grid, UM_THREADS, smem, st>>>(g, a);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

template <int BN, int NSPLIT>
int launch_ns(const GatherGeom& g, const UmmaArgs& a, int nt, bool vec, int Z, cudaStream_t st) {
  if (!vec) return launch_cfg<BN, NSPLIT, false, 1>(g, a, nt, Z, st);
  static const int pf = env_int("SAG_UMMA_PF", 1);
  switch (pf) {
    case 1: return launch_cfg<BN, NSPLIT, true, 1>(g, a, nt, Z, st);
    case 2: return launch_cfg<BN, NSPLIT, true, 2>(g, a, nt, Z, st);
    default: return launch_cfg<BN, NSPLIT, true, 3>(g, a, nt, Z, st);
  }
}

template <int BN>
int launch_bn(const GatherGeom& g, const UmmaArgs& a, int nt, int planes, bool vec, int Z, cudaStream_t st) {
  if (planes == 2) return launch_ns<BN, 3>(g, a, nt, vec, Z, st);
  return launch_ns<BN, 1>(g, a, nt, vec, Z, st);
}

}  // namespace

// ---- host API -----------------------------------------------------------------------------------------------------
void umma_free(UmmaWeights* w) {
  if (w->packed) cudaFree(w->packed);
  if (w->col_off) cudaFree(w->col_off);
  if (w->col_dy) cudaFree(w->col_dy);
  if (w->col_dx) cudaFree(w->col_dx);
  if (w->col_bias) cudaFree(w->col_bias);
  *w = UmmaWeights();
}

int umma_pack_weights(const float* wk, int K, int N, int64_t ldw, int precision, UmmaWeights* out, cudaStream_t st) {
  SAG_REQUIRE(precision == SAG_PREC_BF16 || precision == SAG_PREC_BF16X3, SAG_EUNSUPPORTED,
              "tcgen05 path: precision %d is not built (use bf16 or bf16x3)", precision);
  SAG_REQUIRE(K >= 0 && N > 0, SAG_EINVAL, "umma_pack_weights: bad shape %dx%d", K, N);
  UmmaWeights w;
  w.K = K; w.N = N;
  w.KC = cdiv(K, UM_BK);
  w.BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);   // == tile_width(N)
  w.NT = cdiv(N, w.BN);
  w.planes = precision == SAG_PREC_BF16X3 ? 2 : 1;
  const size_t bytes = (size_t)w.NT * w.KC * w.planes * w.BN * 128;
  if (bytes > 0) {
    SAG_CHECK_CUDA(cudaMalloc(&w.packed, bytes));
    const int64_t total = (int64_t)w.NT * w.KC * w.BN * 8;
    int64_t blocks = cdiv64(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    umma_pack_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(wk, K, N, ldw, w.BN, w.KC, w.NT, w.planes,
                                                              reinterpret_cast<uint8_t*>(w.packed));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(w.packed); set_error("umma_pack_weights: %s", cudaGetErrorString(e)); return SAG_ECUDA; }
  }
  *out = w;
  return SAG_OK;
}

int umma_pack_conv_expanded(const float* w_hwio, int kh, int kw, int cin, int cout, int kw2, int cin2, int precision,
                            UmmaWeights* out, cudaStream_t st) {
  float* wk = nullptr;
  const int64_t total = (int64_t)kh * kw2 * cin2 * cout;
  SAG_CHECK_CUDA(cudaMalloc(&wk, sizeof(float) * (size_t)total));
  int64_t blocks = cdiv64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  expand_hwio_kernel<<<(unsigned)blocks, 256, 0, st>>>(w_hwio, kh, kw, cin, cout, kw2, cin2, wk);
  int r = umma_pack_weights(wk, kh * kw2 * cin2, cout, cout, precision, out, st);
  cudaStreamSynchronize(st);
  cudaFree(wk);
  return r;
}

// Sub-pixel formulation of tf.nn.conv2d_transpose VALID (core.py:139-140): every cell (u, v) of the
// (H+ty-1) x (W+tx-1) grid produces its sh x sw x Cout outputs from ty x tx taps of the input.
int umma_pack_deconv(const float* w_hwoi, const float* bias, int kh, int kw, int cout, int cin, int sh, int sw, int order,
                     int64_t y_sh, int64_t y_sw, int64_t y_sc, int precision, UmmaWeights* out, cudaStream_t st) {
  const int ty = cdiv(kh, sh), tx = cdiv(kw, sw);
  const int K = ty * tx * cin, N = sh * sw * cout;
  SAG_REQUIRE(ty * tx <= kMaxTaps, SAG_EINVAL, "deconv: too many taps");
  float* wk = nullptr;
  SAG_CHECK_CUDA(cudaMalloc(&wk, sizeof(float) * (size_t)K * N));
  {
    const int64_t total = (int64_t)K * N;
    int64_t blocks = cdiv64(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    subpixel_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(w_hwoi, kh, kw, cout, cin, sh, sw, ty, tx, order, wk);
  }
  UmmaWeights w;
  int r = umma_pack_weights(wk, K, N, N, precision, &w, st);
  cudaStreamSynchronize(st);
  cudaFree(wk);
  SAG_TRY(r);
  std::vector<int> off(N);
  std::vector<short> dy(N), dx(N);
  bool vec = true;
  for (int n = 0; n < N; ++n) {
    int py, px, co;
    if (order == 0) { co = n % cout; int ph = n / cout; px = ph % sw; py = ph / sw; }
    else { px = n % sw; int q = n / sw; co = q % cout; py = q / cout; }
    off[n] = (int)(py * y_sh + px * y_sw + co * y_sc);
    dy[n] = (short)py;
    dx[n] = (short)px;
  }
  for (int n = 0; n + 3 < N; n += 4)
    for (int e = 1; e < 4; ++e)
      if (off[n + e] != off[n] + e || dy[n + e] != dy[n] || (order == 0 && dx[n + e] != dx[n])) vec = false;
  // order 1 groups run along px: all four must be valid together, which holds when the output width is a multiple of 4
  // cells wide (checked by the caller through vec4_ok); offsets must keep 16-byte alignment
  if (N % 4 != 0) vec = false;
  for (int n = 0; n < N; n += 4)
    if (off[n] % 4 != 0) vec = false;
  w.vec4 = vec ? 1 : 0;
  cudaError_t e = cudaSuccess;
  if ((e = cudaMalloc(&w.col_off, sizeof(int) * N)) != cudaSuccess || (e = cudaMalloc(&w.col_dy, sizeof(short) * N)) != cudaSuccess ||
      (e = cudaMalloc(&w.col_dx, sizeof(short) * N)) != cudaSuccess || (e = cudaMalloc(&w.col_bias, sizeof(float) * N)) != cudaSuccess) {
    umma_free(&w);
    set_error("umma_pack_deconv: %s", cudaGetErrorString(e));
    return SAG_ECUDA;
  }
  cudaMemcpy(w.col_off, off.data(), sizeof(int) * N, cudaMemcpyHostToDevice);
  cudaMemcpy(w.col_dy, dy.data(), sizeof(short) * N, cudaMemcpyHostToDevice);
  cudaMemcpy(w.col_dx, dx.data(), sizeof(short) * N, cudaMemcpyHostToDevice);
  if (bias != nullptr) {
    expand_bias_kernel<<<cdiv(N, 128), 128, 0, st>>>(bias, cout, sw, N, order, w.col_bias);
  } else {
    cudaMemsetAsync(w.col_bias, 0, sizeof(float) * N, st);
  }
  SAG_CHECK_CUDA(cudaStreamSynchronize(st));
  *out = w;
  return SAG_OK;
}

// Geometry of the sub-pixel GEMM for output rows [row0,row1) of the full transposed-conv output.
int make_deconv_subpixel_geom(GatherGeom* g, int n, int h, int w, int cin, int64_t x_ld, int kh, int kw, int sh, int sw,
                              int row0, int row1, int64_t y_sn, int64_t y_sh, int64_t y_sw, int64_t y_sc, int* oh_lim,
                              int* ow_lim) {
  const int ty = cdiv(kh, sh), tx = cdiv(kw, sw);
  const int OHf = (h - 1) * sh + kh, OWf = (w - 1) * sw + kw;
  if (row1 > OHf) row1 = OHf;
  SAG_REQUIRE(row0 >= 0 && row1 > row0, SAG_EINVAL, "deconv: empty row range");
  const int u0 = row0 / sh, u1 = (row1 - 1) / sh;       // grid rows that own at least one requested output row
  memset(g, 0, sizeof(*g));
  g->N = n; g->H = h; g->W = w; g->Cin = cin; g->x_ld = x_ld;
  g->PH = u1 - u0 + 1;
  g->PW = cdiv(OWf, sw);
  g->isy = 1; g->isx = 1;
  g->oy0 = u0 * sh - row0; g->ox0 = 0; g->osy = sh; g->osx = sw;
  g->y_sn = y_sn; g->y_sh = y_sh; g->y_sw = y_sw; g->y_sc = y_sc;
  g->T = ty * tx;
  for (int a = 0; a < ty; ++a)
    for (int b = 0; b < tx; ++b) {
      int t = a * tx + b;
      g->dy[t] = (short)(u0 - a);
      g->dx[t] = (short)(-b);
      g->widx[t] = (short)t;
    }
  *oh_lim = row1 - row0;
  *ow_lim = OWf;
  return SAG_OK;
}

// Split-K plan: layers whose tile count cannot fill the 148 SMs but whose K loop is long are cut along K; the
// partial accumulators go through `scratch` ([Z][M][n_pad] fp32) and splitk_reduce_kernel finishes them
// (deterministic: fixed summation order).
static int tile_width(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 128); }

int umma_split_k(int K, int N, int64_t M, size_t* scratch_bytes) {
  static const int enabled = env_int("SAG_UMMA_SPLITK", 1);
  const int BN = tile_width(N), NT = cdiv(N, BN), KC = cdiv(K, UM_BK);
  const int64_t tiles = cdiv64(M, UM_BM) * NT;
  int Z = 1;
  if (enabled && tiles > 0 && tiles < 120 && KC >= 8) {
    Z = (int)(296 / tiles);                 // aim at ~2 CTAs worth of work per SM
    if (Z > KC / 4) Z = KC / 4;             // at least 4 chunks per split
    if (Z > 32) Z = 32;
    if (Z < 1) Z = 1;
  }
  if (scratch_bytes) *scratch_bytes = Z > 1 ? sizeof(float) * (size_t)Z * (size_t)M * (size_t)(NT * BN) : 0;
  return Z;
}

int launch_gather_gemm_umma(const float* x, const UmmaWeights& w, float* y, const GatherGeom& g, const Epilogue& ep,
                            int oh_lim, int ow_lim, float* scratch, cudaStream_t st) {
  SAG_REQUIRE(w.packed != nullptr || w.KC == 0, SAG_ESTATE, "tcgen05 path: weights are not packed");
  SAG_REQUIRE(g.T * g.Cin == w.K, SAG_EINVAL, "tcgen05 path: geometry K %d does not match packed K %d", g.T * g.Cin, w.K);
  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  if (M == 0) return SAG_OK;
  SAG_REQUIRE(cdiv64(M, UM_BM) < (1ll << 31), SAG_EINVAL, "tcgen05 path: too many rows");
  UmmaArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.wpacked = reinterpret_cast<const uint8_t*>(w.packed); a.y = y;
  a.relu = ep.relu; a.stat_sum = ep.stat_sum; a.stat_sqs = ep.stat_sqs;
  a.col_off = w.col_off; a.col_dy = w.col_dy; a.col_dx = w.col_dx;
  a.bias = w.col_off != nullptr ? w.col_bias : ep.bias;
  a.oh_lim = oh_lim; a.ow_lim = ow_lim;
  a.Ntot = w.N; a.K = w.K; a.KC = w.KC;
  a.n_pad = w.NT * w.BN;
  const bool aligned_y = (reinterpret_cast<uintptr_t>(y) & 15) == 0;
  if (w.col_off != nullptr) {
    SAG_REQUIRE(ep.stat_sum == nullptr, SAG_EUNSUPPORTED, "tcgen05 path: statistics with mapped outputs");
    a.vec_store = (w.vec4 && aligned_y && g.y_sn % 4 == 0 && g.y_sh % 4 == 0 && (g.y_sw * g.osx) % 4 == 0 && ow_lim % 4 == 0) ? 1 : 0;
  } else {
    a.vec_store = (g.y_sc == 1 && aligned_y && w.N % 4 == 0 && g.y_sn % 4 == 0 && g.y_sh % 4 == 0 && g.y_sw % 4 == 0) ? 1 : 0;
  }
  // vector gather: every 8-element K group is one tap's 8 contiguous, 16-byte aligned floats
  bool vec = (g.Cin % 8 == 0) && ((g.x_ld * g.isx) % 4 == 0) && ((g.x_ld * g.W) % 4 == 0) &&
             ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int t = 0; t < g.T && vec; ++t) vec = (g.dx[t] * g.x_ld) % 4 == 0;
  int Z = scratch != nullptr ? umma_split_k(w.K, w.N, M, nullptr) : 1;
  if (Z > 1) a.partial = scratch;
  int r;
  switch (w.BN) {
    case 32: r = launch_bn<32>(g, a, w.NT, w.planes, vec, Z, st); break;
    case 64: r = launch_bn<64>(g, a, w.NT, w.planes, vec, Z, st); break;
    case 128: r = launch_bn<128>(g, a, w.NT, w.planes, vec, Z, st); break;
    default: set_error("tcgen05 path: unsupported tile width %d", w.BN); return SAG_EINVAL;
  }
  SAG_TRY(r);
  if (Z > 1) {
    const int rows_per_block = 8;
    splitk_reduce_kernel<<<(unsigned)cdiv64(M, rows_per_block), 128, 0, st>>>(g, a, Z, rows_per_block);
    SAG_LAUNCH_CHECK();
  }
  return SAG_OK;
}

}  // namespace sag
