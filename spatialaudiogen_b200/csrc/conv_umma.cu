// tcgen05 gather-GEMM: the tensor-core implementation of every dense contraction on the path -- tfw.conv_2d
// (reference core.py:156-220), tfw.deconv_2d (core.py:96-153, as one sub-pixel GEMM per layer) and
// tfw.fully_connected (core.py:43-93).  im2col-free: the A operand is gathered straight from the NHWC activation.
//
//   C[m, n] = sum_{t, ci} X[pixel(m) + tap t, ci] * Wk[t*Cin + ci, n]          (zero outside the image)
//
// Activations travel between layers as two bf16 planes (hi = bf16(x), lo = bf16(x - hi); same bytes as fp32), written
// once by whichever kernel produces them, so the gather is a pure copy.  Layers whose channel count is a multiple of 64
// (all but the first two audio convolutions) are fed by the TMA unit in im2col mode: one thread issues, per 64-wide K
// chunk, two cp.async.bulk.tensor.4d...im2col copies (hi / lo plane: 128 output pixels x 64 channels at the chunk's
// filter offset, zero outside the image, landing in the swizzled operand layout) + one bulk copy of the packed weight
// tile.  The others gather with 16-byte cp.async (LDGSTS, zero-fill); fp32 sources (stage entry points) take a register
// path that splits on the fly.
//
// Persistent, warp-specialised: one CTA per SM walks the list of (M tile, N tile, K split) work items; tile = 128 x BN.
//   TMA gather      : warp 0 producer (one elected lane), warps 4-11 epilogue (two per TMEM lane quadrant)
//   cp.async gather : warps 0-7 producers (asynchronous mbarrier arrival), warps 8-11 epilogue
//   epilogue        : tcgen05.ld of the accumulator, bias / ReLU, staging tile, coalesced row writes as fp32 / split
//                     bf16 / split-K partials, batch-norm column sums
//   warp 12         : allocates TMEM (two accumulators) and issues tcgen05.mma kind::f16 through one elected lane:
//               SAG_PREC_BF16   : 1 MMA per K step  (A_hi x B_hi)
//               SAG_PREC_BF16X3 : A_hi x B_hi + A_hi x B_lo + A_lo x B_hi -> fp32-grade products; tiles up to 128 wide
//                                 do it with 2 MMAs (A_hi x [B_hi | B_lo] at width 2*BN, A_lo x B_hi), 256-wide with 3
// full/empty mbarrier ring between producers and the MMA issuer, tcgen05.commit releases stages and publishes the
// accumulator; tmem_full/tmem_empty barriers between the MMA issuer and the epilogue, so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Launched with programmatic stream serialisation: the prologue (barriers, TMEM) overlaps the
// previous kernel's tail.  Roofline: tensor pipe; the measured limit today is the shared-memory port (operand reads of
// the 2-3 MMAs per product + stage writes) on narrow tiles and the epilogue's global stores on small-K layers -- see
// DESIGN.md section 3.
#include "model.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace sag {

namespace {

constexpr int UM_BM = 128;                  // rows per tile (UMMA M, one TMEM lane per row)
constexpr int UM_BK = 64;                   // K elements per stage = one 128-byte swizzle atom of bf16
constexpr int UM_PRODUCER_WARPS = 8;         // warps 0-7
constexpr int UM_EPI_WARP0 = 8;              // warps 8-11: epilogue (warp % 4 == TMEM lane quadrant); + warps 4-7 with the TMA producer
constexpr int UM_MMA_WARP = 12;              // warp 12: TMEM allocation + MMA issue (TMA producer: warp 1 -- twelve warps, so 168 registers each)
constexpr int UM_THREADS = 13 * 32;
__host__ __device__ constexpr int um_threads(int src) { return src == 3 ? 12 * 32 : UM_THREADS; }   // (3 = SRC_TMA: warp 0 produces, 1 issues the MMAs, 4-11 epilogue)
constexpr int UM_MAX_N = 1024;               // widest GEMM with batch-norm statistics / per-CTA column sums
constexpr int UM_STAGING_BYTES = UM_BM * (32 * 4 + 16);
constexpr long long UM_ROW_INVALID = (long long)0x8000000000000000ull;   // row beyond M (mapped outputs may have negative bases)
constexpr int UM_A_PLANE = UM_BM * 128;     // bytes of one A plane per stage
constexpr int UM_MAX_STAGES = 8;
constexpr int UM_BAR_BYTES = 256;            // full[8] empty[8] tmem_full[2] tmem_empty[2] tmem slot peer[8]

struct UmmaArgs {
  const void* x;          // fp32 activation, or the hi plane of a split-bf16 activation (src_bf2)
  int64_t x_plane;        // bytes from the hi plane to the lo plane
  const uint8_t* wpacked;
  void* y;                // fp32 output, or the hi plane of a split-bf16 output (out_bf2: 1 = hi only, 2 = hi + lo)
  int64_t y_plane;
  int out_bf2;
  const float* bias;      // indexed by GEMM column (already expanded for mapped outputs) or null
  int relu;
  unsigned long long* stat_sum;      // fixed-point accumulators, two words per column (fx_atomic_add)
  unsigned long long* stat_sqs;
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
  int dense;              // output pixel m sits at element m*y_sw (no row decode)
  int NT, Z;              // N tiles, K splits
  int tma_w0, tma_h0;     // SRC_TMA: smallest tap displacement (= lower corner of the im2col bounding box)
  long long* trace;       // SAG_UMMA_TRACE (debug): per-CTA cycle counters of the three roles
  int dbg;                // SAG_UMMA_EPI_DEBUG (timing experiments, results invalid): 1 no statistics, 2 no TMA store, 4 no staging writes
  int pair_ok;            // host: the tiled weight map exists (CTA pairs fetch their weight blocks through it)
  // how a staged 128 x 32 pass leaves the CTA: 0 = LSU copy-out (any output format / mapping); 1 = one TMA tensor store per
  // pass (dense fp32 rows, or the split-K partials); 2 = bulk stores of contiguous 4 KB runs (sub-pixel transposed conv
  // whose tile is one grid row: the 128 rows x 8 pixels of a column group are 1024 consecutive floats of the output);
  // 3 = the track logits never leave the SM: sigmoid + the 32 -> 9 localization-weighted sums in registers, 9 runs of 4 KB out
  int out_mode;
  int64_t m_pad;          // rows of one split-K partial slab (M rounded up to whole tiles)
  // out_mode 3: mask-gain fusion (Epilogue::gains): column order 2, 256-wide tiles, tile = one grid row of one window
  const float* gain_loc;
  float* gains;
  int64_t gain_plane;
  // stream-K (sk_flags != null, Z == 1): the (tile, K chunk) units are dealt out evenly over the clusters, a cluster's range is
  // cut at tile boundaries; a piece that does not end its tile leaves its raw accumulators in the CTA's slab and raises the
  // CTA's flag, the piece that ends the tile adds the slabs of the pieces before it in its epilogue (fixed order: reproducible)
  int* sk_flags;          // one per CTA, zero between launches (the consumer clears what it has seen)
  float* sk_slab;         // [CTA][128][BN] fp32
};
// im2col tensor maps of the two activation planes (SRC_TMA) + the tiled map of the packed weight image (CTA pairs: 128-byte
// rows, boxes of BN/2 rows); kernel parameter, read by the TMA unit
struct alignas(64) TmaPair { CUtensorMap hi, lo, w, o; };     // o: tiled map of the output (out_mode 1)
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
// im2col tensor copy (NHWC activation, rank 4): 128 output pixels starting at base pixel (w, h, n) -- the hardware walks
// the bounding box of the tensor map with the convolution stride, wrapping rows and images -- x 64 channels from c, at
// filter offset (off_w, off_h); lands as 128 rows of 128 bytes in the SWIZZLE_128B pattern; out-of-image pixels read 0.
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst_smem, const CUtensorMap* tmap, int c, int w, int h, int n,
                                              uint32_t bar, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// CTA-pair variants (.cta_group::2): the copy lands in the executing CTA's shared memory, its bytes complete on an mbarrier
// that may live in the peer CTA (the pair leader's full barrier): no relay between the CTAs
__device__ __forceinline__ void tma_im2col_4d_2sm(uint32_t dst_smem, const CUtensorMap* tmap, int c, int w, int h, int n,
                                                  uint32_t bar_cluster, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_tile_2d_2sm(uint32_t dst_smem, const CUtensorMap* tmap, int x, int y, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
// shared -> global through the TMA unit (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t src_smem, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(src_smem), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Column sums over the 32 lanes of a warp of CPT values per lane (16 or 32): butterfly that halves the live values per
// step, so CPT + (CPT == 16) shuffles instead of 5*CPT.  Returns the total of column ((lane >> 1) & 15) [CPT 16] /
// column lane [CPT 32].
template <int CPT>
__device__ __forceinline__ float warp_column_sums(float (&v)[CPT], int lane) {
  static_assert(CPT == 16 || CPT == 32, "CPT");
  int n = CPT;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      const int h = n >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < CPT / 2; ++i) {
        if (i < h) {
          const float send = upper ? v[i] : v[i + h];
          const float keep = upper ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      n = h;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  return v[0];
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {      // same offset in CTA `rank` of the cluster
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr), "r"(rank));
  return remote;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {      // acquire at cluster scope
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// CTA-pair variants (cta_group::2): the TMEM allocation covers both CTAs, one MMA spans M = 256 (each CTA's own 128-row
// A tile) and reads half of the N rows of B from each CTA's shared memory, commits arrive on both CTAs' barriers
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar) {       // arrives on the barrier at this offset in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
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

// the same without the wait: several loads in flight, then one tmem_ld_wait() before the first use
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// 4 fp32 -> 4 bf16 hi (8 bytes) and 4 bf16 lo
__device__ __forceinline__ void split4(const float* f, uint2& hi, uint2& lo) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 l0 = __floats2bfloat162_rn(f[0] - __low2float(h0), f[1] - __high2float(h0));
  __nv_bfloat162 l1 = __floats2bfloat162_rn(f[2] - __low2float(h1), f[3] - __high2float(h1));
  hi = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  lo = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
}
// store one value of a split-bf16 tensor (element index e from the hi plane base)
__device__ __forceinline__ void store_bf2_1(void* yhi, int64_t plane, int64_t e, float v, int planes) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  reinterpret_cast<__nv_bfloat16*>(yhi)[e] = h;
  if (planes == 2)
    reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(yhi) + plane)[e] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ void store_bf2_4(void* yhi, int64_t plane, int64_t e, const float* v, int planes) {
  uint2 hi, lo;
  split4(v, hi, lo);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(yhi) + e) = hi;
  if (planes == 2) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(yhi) + plane) + e) = lo;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// SRC: 0 = fp32 activations, element-wise gather; 1 = fp32, every 8-element K group lies inside one tap and is
// contiguous + 16-byte aligned (Cin % 8 == 0): vector gather; 2 = split-bf16 planes, same alignment rule: cp.async gather.
// 3 = split-bf16 planes through the TMA unit in im2col mode (Cin % 64 == 0): one thread per CTA feeds the operand ring.
constexpr int SRC_F32 = 0, SRC_F32_VEC = 1, SRC_BF2 = 2, SRC_TMA = 3;

// Persistent: one CTA per SM walks the work list (m tile, n tile, K split) with a static stride; the three roles run
// decoupled through mbarriers, so the operand ring never drains between tiles and the epilogue of tile i overlaps the
// MMAs of tile i+1 (two TMEM accumulators).
// Stream-K work of one cluster: its units [u0, u1) of the tiles * KC (tile, K chunk) units, as n pieces of consecutive tiles.
// Order within the cluster: the piece that leaves its tile unfinished first (its consumer finds the slab ready), whole tiles
// next, the piece that finishes a tile begun by the clusters before it last.  (Host + device: sag_stream_k_schedule runs the
// same code for the CPU tests.)
struct SkRange {
  int64_t u0, u1, t0, U, G, c;
  int KC, n;
  bool rot;
  __host__ __device__ void init(int64_t tiles, int kc, int64_t clusters, int64_t cluster) {
    KC = kc; G = clusters; c = cluster;
    U = tiles * kc;
    u0 = c * U / G;
    u1 = (c + 1) * U / G;
    t0 = 0; n = 0; rot = false;
    if (u1 > u0) {
      t0 = u0 / KC;
      n = (int)((u1 - 1) / KC - t0 + 1);
      rot = (u1 % KC) != 0 && n > 1;
    }
  }
  // j-th piece in execution order: tile, K chunks [kb, ke); produce: the piece does not end its tile (raw accumulators ->
  // slab, raise the flag); fix_first >= 0: the piece ends a tile begun elsewhere -- clusters fix_first .. c-1 hold the rest
  __host__ __device__ void item(int64_t j, int64_t* tile, int* kb, int* ke, bool* produce, int* fix_first) const {
    int64_t i = 0;
    if (n > 1) i = rot ? (j == 0 ? n - 1 : (j == n - 1 ? 0 : j)) : (j == n - 1 ? 0 : j + 1);
    const int64_t t = t0 + i, tb = t * KC;
    *tile = t;
    *kb = (int)((u0 > tb ? u0 : tb) - tb);
    *ke = (int)((u1 < tb + KC ? u1 : tb + KC) - tb);
    *produce = *ke < KC;
    *fix_first = -1;
    if (!*produce && *kb > 0) {
      int64_t f = c - 1;
      while (f > 0 && f * U / G > tb) --f;
      *fix_first = (int)f;
    }
  }
};

// SAG_CONV_REG128 (development): declare 512 threads per block to ptxas, which caps the TMA-fed kernels at 128 registers and leaves
// a quarter of the register file to CTAs of other kernels (batch-norm passes of another lane) on the same SM
#ifdef SAG_CONV_REG128
#define SAG_UM_LB(src) ((src) == 3 ? 512 : um_threads(src))
#define SAG_HL_LB 512
#else
#define SAG_UM_LB(src) um_threads(src)
#define SAG_HL_LB HL_THREADS
#endif
template <int BN, int NSPLIT, int SRC, bool PAIR>
__global__ void __launch_bounds__(SAG_UM_LB(SRC), 1)
gather_gemm_umma_kernel(const __grid_constant__ GatherGeom g, const UmmaArgs a, const __grid_constant__ TmaPair tm) {
  constexpr bool VEC = SRC == SRC_F32_VEC;
  constexpr int PLANES = NSPLIT >= 2 ? 2 : 1;
  constexpr bool TMA_ANY = SRC == SRC_TMA;
  constexpr int EW = TMA_ANY ? 8 : 4;         // epilogue warps (see the epilogue role)
  constexpr bool CONCAT = NSPLIT == 2;        // A_hi x [B_hi | B_lo] as one MMA of width 2*BN, then A_lo x B_hi
  constexpr int B_PLANE = BN * 128;
  // B rows per CTA and stage.  Alone: [B_hi | B_lo] (BN rows each).  CTA pair (each CTA feeds half of the N rows of an
  // MMA, at the same shared-memory offset in both): block X = B_hi (rank 0) / B_lo (rank 1) for the 2*BN-wide MMA of
  // the two-MMA scheme, block Y = this CTA's half of B_hi for the BN-wide one; three-MMA scheme: X / Y = this CTA's
  // half of B_hi / B_lo.
  constexpr int BX_ROWS = !PAIR ? BN : (NSPLIT == 2 ? BN : BN / 2);
  constexpr int BY_ROWS = PLANES == 1 ? 0 : (!PAIR ? BN : BN / 2);
  constexpr int B_BYTES = (BX_ROWS + BY_ROWS) * 128;
  constexpr int STAGE_BYTES = PLANES * UM_A_PLANE + B_BYTES;
  static_assert(!PAIR || SRC == SRC_TMA, "CTA pairs need the TMA producer");
  constexpr int ACC_COLS = CONCAT ? 2 * BN : BN;                 // TMEM columns of one accumulator
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;      // two accumulators
  static_assert(TMEM_COLS <= 512, "accumulators exceed TMEM");
  constexpr uint32_t UM_M = PAIR ? 2 * UM_BM : UM_BM;            // rows of one MMA
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((UM_M >> 4) << 24);
  constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((UM_M >> 4) << 24);

  extern __shared__ __align__(16) uint8_t um_smem[];
  __shared__ float s_sum[UM_MAX_N], s_sqs[UM_MAX_N];     // batch-norm partial sums of this CTA, by absolute column
  __shared__ long long s_yoff[UM_BM];                    // per-row output element offset of the tile in the epilogue
  __shared__ int s_oy[UM_BM], s_ox[UM_BM];
  __shared__ float s_part[2][8][64];          // per-epilogue-warp column sums / sums of squares of a pass (by pass parity)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bars = smem_u32(um_smem);
  const uint32_t bar_full = bars, bar_empty = bars + 8 * UM_MAX_STAGES;
  const uint32_t bar_tfull = bars + 16 * UM_MAX_STAGES, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t stile = (bars + UM_BAR_BYTES + 1023u) & ~1023u;             // epilogue staging tile (128 x 144 B; 1024-byte aligned: swizzled TMA stores)
  const uint32_t tiles = (stile + UM_STAGING_BYTES + 1023u) & ~1023u;        // operand stage ring
  const int S = a.stages;

  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  constexpr int CL = PAIR ? 2 : 1;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int MT = (int)((M + UM_BM - 1) / UM_BM);
  const int MG = (MT + CL - 1) / CL;             // groups of CL adjacent M tiles: one per CTA of a cluster
  const int NT = a.NT, Z = a.Z;
  const int64_t n_work = (int64_t)MG * NT * Z;   // work items per cluster
  const int64_t wk0 = blockIdx.x / CL, wk_step = gridDim.x / CL;
  const bool stats = a.stat_sum != nullptr && a.partial == nullptr;

  // ---- one-time setup ----
  if (a.trace && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.trace[blockIdx.x * 16 + 13] = (long long)t;            // (trace: kernel entry of this CTA, ns)
  }
  constexpr int MMA_WARP = TMA_ANY ? 1 : UM_MMA_WARP;
  constexpr int NTHREADS = um_threads(SRC);
  if (stats) for (int i = tid; i < UM_MAX_N; i += NTHREADS) { s_sum[i] = 0.f; s_sqs[i] = 0.f; }
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        // every producer thread + the expect_tx arrival of the B copy; TMA gather: the expect_tx arrival alone
        mbar_init(bar_full + 8 * s, TMA_ANY ? 1 : UM_PRODUCER_WARPS * 32 + 1);
        mbar_init(bar_empty + 8 * s, 1);                       // one tcgen05.commit
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_tfull + 8 * b, 1);                       // one tcgen05.commit per tile
        mbar_init(bar_tempty + 8 * b, PAIR ? 2 * EW : EW);     // the epilogue warps (of both CTAs) have drained the accumulator
      }
      fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) tmem_alloc2(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();                              // peers' barriers exist before anyone arrives remotely
  else __syncthreads();
  tc_fence_after();
  // everything above (barriers, TMEM, zeroed statistics) is independent of earlier kernels: with programmatic dependent
  // launch it overlaps the tail of the previous kernel; from here on this kernel reads / overwrites global tensors
  pdl_prologue();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // work item -> (m tile, n tile, split): m fastest, so neighbouring CTAs share the weight tile in L2; the CTAs of a
  // cluster take adjacent m tiles of the same (n tile, split) and walk their K chunks in lockstep (mt may be >= MT
  // for the last group: that CTA then runs an all-padding tile to keep the lockstep)
  struct WorkItem {
    int mt, nt, z, kc_begin, kc_end;
    bool produce;      // stream-K piece that does not end its tile: raw accumulators -> slab, raise the flag
    int fix_first;     // stream-K piece that ends a tile begun elsewhere: first cluster holding an earlier piece (-1: none)
  };
  const bool sk = a.sk_flags != nullptr;
  const int64_t sk_c = blockIdx.x / CL;
  SkRange skr;
  skr.init(sk ? (int64_t)MG * NT : 0, a.KC, gridDim.x / CL, sk_c);
  const int64_t n_items = sk ? (int64_t)skr.n : (wk0 < n_work ? (n_work - wk0 + wk_step - 1) / wk_step : 0);
  auto item_at = [&](int64_t j, WorkItem& it) {
    int64_t wk;
    it.produce = false;
    it.fix_first = -1;
    if (!sk) wk = wk0 + j * wk_step;
    else skr.item(j, &wk, &it.kc_begin, &it.kc_end, &it.produce, &it.fix_first);
    it.mt = (int)(wk % MG) * CL + (int)crank;
    const int64_t r = wk / MG;
    it.nt = (int)(r % NT);
    it.z = (int)(r / NT);
    if (!sk) {
      it.kc_begin = (int)((int64_t)it.z * a.KC / Z);
      it.kc_end = (int)((int64_t)(it.z + 1) * a.KC / Z);
    }
  };

  if (TMA_ANY && warp == 0) {
    // ================================ producer (TMA im2col): warp 0, one elected lane (warps 2-3 idle) ================================
    {
      int stage = 0;
      uint32_t phase = 0;
      long long tr_wait = 0, tr_t0 = clock64(), tr_chunks = 0;
      const uint32_t lead_full = PAIR ? map_to_cta(bar_full, 0) : bar_full;      // the pair leader's full barriers (cluster address)
      for (int64_t wj = 0; wj < n_items; ++wj) {
        WorkItem wi;
        item_at(wj, wi);
        const int mt = wi.mt, nt = wi.nt, kc_begin = wi.kc_begin, kc_end = wi.kc_end;
        // base input pixel of the tile's first row (the padding tile of an odd pair reloads the last tile; its rows
        // are never stored)
        const uint32_t mu = (uint32_t)((int64_t)(mt < MT ? mt : MT - 1) * UM_BM);
        const uint32_t q = mu / (uint32_t)g.PW, j = mu - q * (uint32_t)g.PW;
        const uint32_t n = q / (uint32_t)g.PH, i = q - n * (uint32_t)g.PH;
        const int cw = (int)j * g.isx + a.tma_w0, ch = (int)i * g.isy + a.tma_h0;
#pragma unroll 1
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          const long long tr_w0 = a.trace ? clock64() : 0;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (a.trace) { tr_wait += clock64() - tr_w0; ++tr_chunks; }
          if (elect_one()) {
            const int kk = kc * UM_BK;
            const uint32_t st_base = tiles + (uint32_t)stage * STAGE_BYTES;
            const int t = kk / g.Cin;
            const int ci0 = kk - t * g.Cin;
            const uint16_t ow = (uint16_t)(g.dx[t] - a.tma_w0), oh = (uint16_t)(g.dy[t] - a.tma_h0);
            const uint32_t bx = st_base + PLANES * UM_A_PLANE;
            if (!PAIR) {
              const uint32_t bar = bar_full + 8 * stage;
              mbar_arrive_expect_tx(bar, PLANES * UM_A_PLANE + B_BYTES);
              tma_im2col_4d(st_base, &tm.hi, ci0, cw, ch, (int)n, bar, ow, oh);
              if (PLANES == 2) tma_im2col_4d(st_base + UM_A_PLANE, &tm.lo, ci0, cw, ch, (int)n, bar, ow, oh);
              const uint8_t* wsrc = a.wpacked + ((size_t)nt * a.KC + kc) * (size_t)(PLANES * B_PLANE);
              bulk_g2s(bx, wsrc, B_BYTES, bar);
            } else {
              // CTA pair: this CTA's operands land in its own shared memory, their bytes complete on the LEADER's full barrier.
              // The leader's producer makes the stage's single arrival and expects the bytes of BOTH CTAs; the peer only issues
              // its copies (no cross-CTA arrive: a release at cluster scope per stage cost ~900 clk in the producer loop).  Bytes
              // of the peer may land before the leader's expect_tx: the transaction count is signed, the phase cannot complete
              // before the leader's arrival, and the peer cannot be a whole phase ahead (it waits on its own empty barrier,
              // released by the same multicast commit as the leader's).
              const uint32_t bar = lead_full + 8 * stage;
              if (crank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * (PLANES * UM_A_PLANE + B_BYTES));
              tma_im2col_4d_2sm(st_base, &tm.hi, ci0, cw, ch, (int)n, bar, ow, oh);
              if (PLANES == 2) tma_im2col_4d_2sm(st_base + UM_A_PLANE, &tm.lo, ci0, cw, ch, (int)n, bar, ow, oh);
              // packed weight image as rows of 128 bytes: tile (nt, kc) starts at row (nt*KC + kc)*PLANES*BN; boxes of BN/2 rows
              const int wrow = (nt * a.KC + kc) * (PLANES * BN);
              constexpr int HB = BN / 2;
              if (NSPLIT == 2) {
                // block X = B_hi (rank 0) / B_lo (rank 1), block Y = my half of B_hi
                tma_tile_2d_2sm(bx, &tm.w, 0, wrow + (int)crank * BN, bar);
                tma_tile_2d_2sm(bx + HB * 128, &tm.w, 0, wrow + (int)crank * BN + HB, bar);
                tma_tile_2d_2sm(bx + BX_ROWS * 128, &tm.w, 0, wrow + (int)crank * HB, bar);
              } else {
                tma_tile_2d_2sm(bx, &tm.w, 0, wrow + (int)crank * HB, bar);                              // my half of B_hi
                if (PLANES == 2) tma_tile_2d_2sm(bx + BX_ROWS * 128, &tm.w, 0, wrow + BN + (int)crank * HB, bar);   // my half of B_lo
              }
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      if (a.trace && lane == 0) {
        a.trace[blockIdx.x * 16 + 2] = tr_wait;
        a.trace[blockIdx.x * 16 + 3] = clock64() - tr_t0;
        a.trace[blockIdx.x * 16 + 6] = tr_chunks;
      }
    }
  } else if (!TMA_ANY && warp < UM_PRODUCER_WARPS) {
    // ================================ producers ================================
    const int jchunk = tid & 7;                 // which 8-element (16-byte bf16) group of the 64-wide K chunk
    const int64_t x_row = g.x_row != 0 ? g.x_row : (int64_t)g.W * g.x_ld;      // elements between image rows
    int stage = 0;
    uint32_t phase = 0;
    long long tr_wait = 0, tr_t0 = clock64(), tr_chunks = 0;
    for (int64_t wj = 0; wj < n_items; ++wj) {
      WorkItem wi;
      item_at(wj, wi);
      const int mt = wi.mt, nt = wi.nt, kc_begin = wi.kc_begin, kc_end = wi.kc_end;
      const int64_t m0 = (int64_t)mt * UM_BM;
      int iy0[4], ix0[4];
      int64_t img[4];                           // element offset of the row's image
      bool rok[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 32 + (tid >> 3);
        const int64_t m = m0 + r;
        rok[it] = m < M;
        iy0[it] = 0; ix0[it] = 0; img[it] = 0;
        if (rok[it]) {                          // M < 2^31 (checked on the host): 32-bit decode
          const uint32_t mu = (uint32_t)m;
          const uint32_t q = mu / (uint32_t)g.PW, j = mu - q * (uint32_t)g.PW;
          const uint32_t n = q / (uint32_t)g.PH, i = q - n * (uint32_t)g.PH;
          iy0[it] = (int)i * g.isy;
          ix0[it] = (int)j * g.isx;
          img[it] = (int64_t)n * g.H * x_row;
        }
      }
#pragma unroll 1
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        const int kk = kc * UM_BK + jchunk * 8;
        const int t = kk / g.Cin;
        const int ci0 = kk - t * g.Cin;
        float f[4][8];
        if (SRC == SRC_BF2) {
          // nothing to stage in registers: the copies go straight to shared memory below
        } else if (VEC) {
          const bool kok = kk < a.K;
          const int dy = kok ? g.dy[t] : 0, dx = kok ? g.dx[t] : 0;
          const float* xf = reinterpret_cast<const float*>(a.x);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int iy = iy0[it] + dy, ix = ix0[it] + dx;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (kok && rok[it] && (unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W) {
              const float4* p = reinterpret_cast<const float4*>(xf + img[it] + ((int64_t)iy * x_row + (int64_t)ix * g.x_ld) + ci0);
              v0 = __ldg(p);
              v1 = __ldg(p + 1);
            }
            f[it][0] = v0.x; f[it][1] = v0.y; f[it][2] = v0.z; f[it][3] = v0.w;
            f[it][4] = v1.x; f[it][5] = v1.y; f[it][6] = v1.z; f[it][7] = v1.w;
          }
        } else {
          const float* xf = reinterpret_cast<const float*>(a.x);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            int tt = t, ci = ci0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = 0.f;
              if (rok[it] && kk + e < a.K) {
                const int iy = iy0[it] + g.dy[tt], ix = ix0[it] + g.dx[tt];
                if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W)
                  v = __ldg(xf + img[it] + ((int64_t)iy * x_row + (int64_t)ix * g.x_ld) + ci);
              }
              f[it][e] = v;
              if (++ci == g.Cin) { ci = 0; ++tt; }
            }
          }
        }
        // claim the stage: the MMAs that read it last time round have completed; start the weight copy
        const long long tr_w0 = a.trace ? clock64() : 0;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (a.trace) { tr_wait += clock64() - tr_w0; ++tr_chunks; }
        const uint32_t st_base = tiles + (uint32_t)stage * STAGE_BYTES;
        if (tid == 0) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, PLANES * B_PLANE);
          const uint8_t* wsrc = a.wpacked + ((size_t)nt * a.KC + kc) * (size_t)(PLANES * B_PLANE);
          bulk_g2s(st_base + PLANES * UM_A_PLANE, wsrc, PLANES * B_PLANE, bar_full + 8 * stage);
        }
        if (SRC == SRC_BF2) {
          const char* xhi = reinterpret_cast<const char*>(a.x);
          const bool kok = kk < a.K;
          const int dy = kok ? g.dy[t] : 0, dx = kok ? g.dx[t] : 0;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 32 + (tid >> 3);
            const uint32_t off = (uint32_t)r * 128u + (uint32_t)((jchunk ^ (r & 7)) << 4);
            const int iy = iy0[it] + dy, ix = ix0[it] + dx;
            const bool ok = kok && rok[it] && (unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W;
            const char* src = ok ? xhi + 2 * (img[it] + ((int64_t)iy * x_row + (int64_t)ix * g.x_ld) + ci0) : xhi;
            cp_async16(st_base + off, src, ok ? 16u : 0u);
            if (PLANES == 2) cp_async16(st_base + UM_A_PLANE + off, src + (ok ? a.x_plane : 0), ok ? 16u : 0u);
          }
          // asynchronous publication: the full barrier gets this thread's arrival when its copies have landed -- no
          // wait and no fence here (a MEMBAR in the producer would wait for every copy in flight and serialise the
          // ring); the MMA thread orders the generic-proxy writes against the tensor core after its barrier wait
          cp_async_arrive_noinc(bar_full + 8 * stage);
        } else {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 32 + (tid >> 3);
            const uint32_t off = (uint32_t)r * 128u + (uint32_t)((jchunk ^ (r & 7)) << 4);
            uint4 hi, lo;
            split8(f[it], hi, lo);
            st_shared_v4(st_base + off, hi);
            if (PLANES == 2) st_shared_v4(st_base + UM_A_PLANE + off, lo);
          }
          fence_proxy_async();                     // st.shared of this thread -> visible to the tensor core
          mbar_arrive(bar_full + 8 * stage);
        }
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
    if (a.trace && tid == 0) {
      a.trace[blockIdx.x * 16 + 2] = tr_wait;
      a.trace[blockIdx.x * 16 + 3] = clock64() - tr_t0;
      a.trace[blockIdx.x * 16 + 6] = tr_chunks;
    }
  } else if (warp == MMA_WARP) {
    // ================================ MMA issuer ================================
    // The whole warp walks the loop (warp-uniform control flow keeps the shared-memory descriptors in uniform
    // registers); one elected lane issues the tcgen05 instructions.  In a CTA pair only the leader (rank 0) issues:
    // its MMAs span both CTAs' A tiles and accumulators.
    if (!PAIR || crank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int64_t it_local = 0;
      long long tr_wait = 0, tr_wacc = 0, tr_t0 = clock64();
      constexpr uint64_t DESC_HI = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);   // LBO, SBO, version, SW128
      for (int64_t wj = 0; wj < n_items; ++wj, ++it_local) {
        WorkItem wi;
        item_at(wj, wi);
        const int kc_begin = wi.kc_begin, kc_end = wi.kc_end;
        const int b = (int)(it_local & 1);
        const uint32_t use = (uint32_t)(it_local >> 1);
        const long long tr_a0 = a.trace ? clock64() : 0;
        mbar_wait(bar_tempty + 8 * b, (use & 1) ^ 1);      // the epilogue has drained this accumulator
        if (a.trace) tr_wacc += clock64() - tr_a0;
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(b * ACC_COLS);
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          const long long tr_w0 = a.trace ? clock64() : 0;
          mbar_wait(bar_full + 8 * stage, phase);
          if (a.trace) tr_wait += clock64() - tr_w0;
          if (SRC == SRC_BF2) fence_proxy_async();   // cp.async (generic proxy) writes observed through the barrier
          tc_fence_after();
          const uint32_t st_base = tiles + (uint32_t)stage * STAGE_BYTES;
          // descriptor low words (address >> 4) of the four operand planes; +2 per 32-byte K step inside the atom
          const uint64_t da_hi = DESC_HI | (uint64_t)((st_base & 0x3FFFFu) >> 4);
          const uint64_t da_lo = DESC_HI | (uint64_t)(((st_base + UM_A_PLANE) & 0x3FFFFu) >> 4);
          const uint64_t db_hi = DESC_HI | (uint64_t)(((st_base + PLANES * UM_A_PLANE) & 0x3FFFFu) >> 4);
          // second B block of the stage: B_lo (alone) / block Y (pair: see BX_ROWS)
          const uint64_t db_lo = DESC_HI | (uint64_t)(((st_base + PLANES * UM_A_PLANE + BX_ROWS * 128) & 0x3FFFFu) >> 4);
          if (elect_one()) {
            auto mma = [&](uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
              if (PAIR) umma_bf16_2(tmem_acc, da, db, idesc, acc);
              else umma_bf16(tmem_acc, da, db, idesc, acc);
            };
#pragma unroll
            for (int k4 = 0; k4 < UM_BK / 16; ++k4) {
              const uint64_t xa_hi = da_hi + 2 * k4, xa_lo = da_lo + 2 * k4;   // 32 bytes further inside the swizzle atom
              const uint32_t first = (kc > kc_begin || k4 > 0) ? 1u : 0u;
              if (CONCAT) {
                // one 2*BN-wide MMA yields [A_hi.B_hi | A_hi.B_lo] in adjacent accumulator blocks (summed by the
                // epilogue): alone the lo plane of B follows its hi plane in the stage, in a pair rank 0 holds B_hi and
                // rank 1 B_lo in block X; A_lo.B_hi lands on the first block (pair: the two halves of B_hi in block Y)
                mma(xa_hi, db_hi + 2 * k4, IDESC2, first);
                mma(xa_lo, (PAIR ? db_lo : db_hi) + 2 * k4, IDESC, 1u);
              } else {
                mma(xa_hi, db_hi + 2 * k4, IDESC, first);
                if (NSPLIT == 3) {
                  mma(xa_lo, db_hi + 2 * k4, IDESC, 1u);
                  mma(xa_hi, db_lo + 2 * k4, IDESC, 1u);
                }
              }
            }
            if (PAIR) {
              umma_commit2(bar_empty + 8 * stage);     // frees the stage in both CTAs
              if (kc + 1 == kc_end) umma_commit2(bar_tfull + 8 * b);
            } else {
              umma_commit(bar_empty + 8 * stage);      // frees the stage once these MMAs have read it
              if (kc + 1 == kc_end) umma_commit(bar_tfull + 8 * b);   // accumulator complete
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      if (a.trace && lane == 0) {
        a.trace[blockIdx.x * 16 + 0] = tr_wait;
        a.trace[blockIdx.x * 16 + 1] = clock64() - tr_t0;
        a.trace[blockIdx.x * 16 + 7] = tr_wacc;
      }
    }
  } else if ((!TMA_ANY && warp >= UM_EPI_WARP0 && warp < UM_MMA_WARP) || (TMA_ANY && warp >= 4)) {
    // ================================ epilogue ================================
    // EW warps: 4 (one per TMEM lane quadrant) next to the cp.async producers; 8 when the TMA unit gathers and warps 4-7
    // are free: two warps per quadrant, each draining half of the columns of a pass.  The pass is latency bound
    // (~600 dependent instructions per warp), so the second warp per scheduler nearly halves it.
    const int q = warp & 3;                        // TMEM lane quadrant of this warp (hardware rule: warp % 4)
    const int half = (EW == 8 && warp < UM_EPI_WARP0) ? 1 : 0;      // which CPT-column half of a pass this warp drains
    const int ew = half * 4 + q;                   // epilogue warp index 0..EW-1
    const int et = ew * 32 + lane;                 // epilogue thread index 0..EW*32-1
    const int row = q * 32 + lane;                 // tile row (= TMEM lane) this thread drains
    constexpr int CPT = 128 / EW;                  // columns per thread per pass: 32 or 16
    constexpr int RD = 32 / EW;                    // copy-out rows per thread per pass: 8 or 4
    constexpr int RPP = 128 / EW;                  // statistics: rows per partial sum
    constexpr uint32_t PITCH = 32 * 4 + 16;        // 32 fp32 columns per pass; +16 B keeps 16-byte row stores conflict-free
    const bool split = a.partial != nullptr;       // K split: raw accumulators go to the partial region first
    float* yf = reinterpret_cast<float*>(a.y);
    const int lr = lane >> 3, lc = (lane & 7) * 4; // copy-out role: a warp instruction covers 4 rows x 128 contiguous bytes
    int64_t it_local = 0;
    long long tr_wait = 0, tr_t0 = clock64();
    long long tr_p[5] = {0, 0, 0, 0, 0};
    // 64-wide tiles with one N tile (conv1): a thread meets the same 2 x CPT columns in every tile, so the batch-norm sums stay in
    // its registers over all tiles of the CTA and cross the warp once, after the last tile (the butterfly per pass was 20 % of
    // conv1's epilogue, which paces that layer)
    constexpr int DEFER_P = BN == 64 ? 2 : 1;
    const bool defer_stats = BN == 64 && NT == 1 && stats && a.out_mode == 1 && a.partial == nullptr && !sk;
    float racc[DEFER_P][CPT], rsq[DEFER_P][CPT];
#pragma unroll
    for (int p = 0; p < DEFER_P; ++p)
#pragma unroll
      for (int e = 0; e < CPT; ++e) { racc[p][e] = 0.f; rsq[p][e] = 0.f; }
    for (int64_t wj = 0; wj < n_items; ++wj, ++it_local) {
      WorkItem wi;
      item_at(wj, wi);
      const int mt = wi.mt, nt = wi.nt, z = wi.z;
      // stream-K slabs hold a tile as [pass][4-column group][epilogue thread] float4s: producer and consumer are the same thread
      // index of two CTAs, so both sides move 512 contiguous bytes per warp instruction
      const bool rawp = wi.produce;                // piece bound for this CTA's slab
      float4* const my_slab = reinterpret_cast<float4*>(a.sk_slab) + (size_t)blockIdx.x * (UM_BM * BN / 4);
      const int64_t m0 = (int64_t)mt * UM_BM;
      const int n_base = nt * BN;
      const int b = (int)(it_local & 1);
      const uint32_t use = (uint32_t)(it_local >> 1);
      const int rows_valid = (int)((M - m0) < UM_BM ? ((M - m0) > 0 ? (M - m0) : 0) : UM_BM);   // valid rows come first
      if (et < UM_BM) {                            // where row `et` of the tile goes in the output tensor
        const int64_t m = m0 + et;
        long long yo = UM_ROW_INVALID;
        int oy = 0, ox = 0;
        if (m < M) {
          if (a.dense) {
            yo = m * g.y_sw;
          } else {
            const uint32_t mu = (uint32_t)m;
            const uint32_t qq = mu / (uint32_t)g.PW, j = mu - qq * (uint32_t)g.PW;
            const uint32_t n = qq / (uint32_t)g.PH, i = qq - n * (uint32_t)g.PH;
            oy = g.oy0 + (int)i * g.osy;
            ox = g.ox0 + (int)j * g.osx;
            yo = (int64_t)n * g.y_sn + (int64_t)oy * g.y_sh + (int64_t)ox * g.y_sw;
          }
        }
        s_yoff[et] = yo; s_oy[et] = oy; s_ox[et] = ox;
      }

      // copy-out of the staged 128 x 32 pass (+ batch-norm column sums): raw -> this split's partial accumulators,
      // else the layer output in its final format.  Ends with a barrier: staging tile and row table are reusable.
      auto emit_pass = [&](const int c0, const bool raw) {
        long long tp2 = a.trace ? clock64() : 0;
        const int n = n_base + c0 + lc;
        const int ncols = raw ? a.n_pad : a.Ntot;
        if (n < ncols) {
          if (raw || (a.dense && a.vec_store)) {
            // fast path: row rr of the tile lands at a fixed stride; all loads first, then all stores
            float4 w4[RD];
#pragma unroll
            for (int rd = 0; rd < RD; ++rd) {
              const int rr = rd * (EW * 4) + ew * 4 + lr;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w4[rd].x), "=f"(w4[rd].y), "=f"(w4[rd].z), "=f"(w4[rd].w)
                           : "r"(stile + (uint32_t)rr * PITCH + (uint32_t)lc * 4u));
            }
            const int64_t rstride = raw ? (int64_t)a.n_pad : g.y_sw;
            const int64_t base = (raw ? ((int64_t)z * a.m_pad + m0) * a.n_pad : m0 * g.y_sw) + n;
#pragma unroll
            for (int rd = 0; rd < RD; ++rd) {
              const int rr = rd * (EW * 4) + ew * 4 + lr;
              if (rr >= rows_valid) continue;
              const int64_t eoff = base + (int64_t)rr * rstride;
              if (raw) {
                __stcg(reinterpret_cast<float4*>(a.partial + eoff), w4[rd]);
              } else if (a.out_bf2) {
                const float t4[4] = {w4[rd].x, w4[rd].y, w4[rd].z, w4[rd].w};
                store_bf2_4(a.y, a.y_plane, eoff, t4, a.out_bf2);
              } else {
                *reinterpret_cast<float4*>(yf + eoff) = w4[rd];
              }
            }
          } else if (a.col_off != nullptr && a.vec_store) {
            // mapped output (sub-pixel transposed conv), groups of 4 columns contiguous: row table and values of all
            // rows are loaded before the first store
            const int c_off = __ldg(a.col_off + n), c_dy = __ldg(a.col_dy + n), c_dx = __ldg(a.col_dx + n);
            float4 w4[RD];
            long long yo[RD];
            bool ok[RD];
#pragma unroll
            for (int rd = 0; rd < RD; ++rd) {
              const int rr = rd * (EW * 4) + ew * 4 + lr;
              yo[rd] = s_yoff[rr];
              ok[rd] = yo[rd] != UM_ROW_INVALID && (unsigned)(s_oy[rr] + c_dy) < (unsigned)a.oh_lim &&
                       (unsigned)(s_ox[rr] + c_dx) < (unsigned)a.ow_lim;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w4[rd].x), "=f"(w4[rd].y), "=f"(w4[rd].z), "=f"(w4[rd].w)
                           : "r"(stile + (uint32_t)rr * PITCH + (uint32_t)lc * 4u));
            }
#pragma unroll
            for (int rd = 0; rd < RD; ++rd) {
              if (!ok[rd]) continue;
              const int64_t eoff = yo[rd] + c_off;
              if (a.out_bf2) {
                const float t4[4] = {w4[rd].x, w4[rd].y, w4[rd].z, w4[rd].w};
                store_bf2_4(a.y, a.y_plane, eoff, t4, a.out_bf2);
              } else {
                *reinterpret_cast<float4*>(yf + eoff) = w4[rd];
              }
            }
          } else {
            // column map of this thread's 4 columns (fixed for the whole pass)
            int c_off[4] = {0, 0, 0, 0}, c_dy[4] = {0, 0, 0, 0}, c_dx[4] = {0, 0, 0, 0};
            if (a.col_off != nullptr) {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (n + e < a.Ntot) { c_off[e] = __ldg(a.col_off + n + e); c_dy[e] = __ldg(a.col_dy + n + e); c_dx[e] = __ldg(a.col_dx + n + e); }
            }
#pragma unroll 4
            for (int rd = 0; rd < RD; ++rd) {
              const int rr = rd * (EW * 4) + ew * 4 + lr;
              const long long yo = s_yoff[rr];
              if (yo == UM_ROW_INVALID) continue;
              float4 w4;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w4.x), "=f"(w4.y), "=f"(w4.z), "=f"(w4.w)
                           : "r"(stile + (uint32_t)rr * PITCH + (uint32_t)lc * 4u));
              const float t4[4] = {w4.x, w4.y, w4.z, w4.w};
              if (a.col_off == nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (n + e >= a.Ntot) continue;
                  const int64_t eoff = yo + (int64_t)(n + e) * g.y_sc;
                  if (a.out_bf2) store_bf2_1(a.y, a.y_plane, eoff, t4[e], a.out_bf2);
                  else yf[eoff] = t4[e];
                }
              } else {
                const int oy = s_oy[rr], ox = s_ox[rr];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (n + e >= a.Ntot) continue;
                  if (!((unsigned)(oy + c_dy[e]) < (unsigned)a.oh_lim && (unsigned)(ox + c_dx[e]) < (unsigned)a.ow_lim)) continue;
                  const int64_t eoff = yo + c_off[e];
                  if (a.out_bf2) store_bf2_1(a.y, a.y_plane, eoff, t4[e], a.out_bf2);
                  else yf[eoff] = t4[e];
                }
              }
            }
          }
        }
        long long tp3 = a.trace ? clock64() : 0;
        const bool st_pass = stats && !raw;
        if (st_pass) {                             // column sums of the staged pass: EW threads per column, RPP rows each
          const int c = et & 31, part = et >> 5;
          float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < RPP; ++i) {
            const int rr = part + EW * i;
            float x;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(stile + (uint32_t)rr * PITCH + (uint32_t)c * 4u));
            if (rr >= rows_valid) x = 0.f;
            cs[i & 3] += x;
            cq[i & 3] = fmaf(x, x, cq[i & 3]);
          }
          // per-warp partials; the first epilogue warp folds them into the CTA sums after the barrier below, in a fixed
          // order (no atomics: the per-CTA statistics are run-to-run reproducible)
          s_part[0][part][c] = (cs[0] + cs[1]) + (cs[2] + cs[3]);
          s_part[0][part][32 + c] = (cq[0] + cq[1]) + (cq[2] + cq[3]);
        }
        long long tp4 = a.trace ? clock64() : 0;
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");   // staging tile and row table reusable
        if (st_pass && et < 32 && n_base + c0 + et < a.Ntot) {
          // (the next pass rewrites s_part only after its own first barrier, which this warp has not reached yet)
          float ts = 0.f, tq = 0.f;
#pragma unroll
          for (int p = 0; p < EW; ++p) { ts += s_part[0][p][et]; tq += s_part[0][p][32 + et]; }
          s_sum[n_base + c0 + et] += ts;
          s_sqs[n_base + c0 + et] += tq;
        }
        if (a.trace) {
          const long long tp5 = clock64();
          tr_p[1] += tp3 - tp2; tr_p[2] += tp4 - tp3; tr_p[3] += tp5 - tp4; tr_p[4] += 1;
        }
      };

      const long long tr_w0 = a.trace ? clock64() : 0;
      mbar_wait(bar_tfull + 8 * b, use & 1);
      if (a.trace) tr_wait += clock64() - tr_w0;
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + (uint32_t)(b * ACC_COLS) + ((uint32_t)(q * 32) << 16);
      if (wi.fix_first >= 0) {
        // the pieces before this one (clusters fix_first .. sk_c - 1, same CTA rank): wait for their slabs, clear the flags for
        // the next launch (each flag has this one consumer)
        if (et == 0) {
          for (int c = wi.fix_first; c < (int)sk_c; ++c) {
            int* f = a.sk_flags + c * CL + (int)crank;
            int seen;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
              if (seen == 0) __nanosleep(64);
            } while (seen == 0);
            asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(0) : "memory");
          }
        }
        const long long tq0 = a.trace ? clock64() : 0;
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (a.trace) tr_p[2] += clock64() - tq0;     // (trace: time blocked on the flags)
      }
      if (rawp) {
        // ---- stream-K piece that leaves its tile unfinished: raw accumulators -> slab, 64 columns (all TMEM loads in flight) at a time
        const long long tq0 = a.trace ? clock64() : 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
          float v[2][CPT];
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int e = 0; e < CPT; e += 16) tmem_ld16_nowait(tmem_row + (uint32_t)(c0 + 32 * p + half * CPT + e), v[p] + e);
          if (CONCAT) {
            float u[2][CPT];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
              for (int e = 0; e < CPT; e += 16) tmem_ld16_nowait(tmem_row + (uint32_t)(BN + c0 + 32 * p + half * CPT + e), u[p] + e);
            tmem_ld_wait();
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
              for (int e = 0; e < CPT; ++e) v[p][e] += u[p][e];
          } else {
            tmem_ld_wait();
          }
          if (c0 + 64 >= BN) {                     // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(bar_tempty + 8 * b, 0); else mbar_arrive(bar_tempty + 8 * b); }
          }
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int e = 0; e < CPT; e += 4)
              __stcg(my_slab + ((size_t)((c0 >> 5) + p) * (CPT / 4) + (e >> 2)) * (EW * 32) + et, make_float4(v[p][e], v[p][e + 1], v[p][e + 2], v[p][e + 3]));
        }
        // the slab is complete: every thread's stores are ordered before the flag (fence by the writers, barrier, release store)
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (et == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.sk_flags + blockIdx.x), "r"(1) : "memory");
        if (a.trace) tr_p[1] += clock64() - tq0;     // (trace: raw pieces)
        continue;
      }
      if constexpr (BN == 256 && EW == 8 && NSPLIT != 2) {
        if (a.out_mode == 3) {
          // ---- mask-gain fusion (model.py:334 sigmoid mask, :424-432 mixing, folded by linearity like mask_gains_kernel in
          // fft.cu): the tile is grid row (window nimg, row irow) x sub-pixel row py = nt; column order 2 hands this thread
          // (tile row = frequency cell j, half) the 32 track logits of sub-pixels px = 4*half + pp, pp = 0..3, two passes each.
          // G[gi] = sum_k w[gi][k] * sigmoid(logit_k) in track order -- the same fmaf chain as mask_gains_kernel, so the gains
          // are bit-identical to the two-kernel path.  36 accumulators per thread; the logits are never stored.
          float* s_w = &s_part[0][0][0];              // [32 tracks][12]: the 9 weights of a track side by side (no statistics here)
          const int gi_row = mt < MT ? mt : MT - 1;   // (the all-padding tile of an odd CTA pair reads valid weights, stores nothing)
          const int nimg = gi_row / g.PH, irow = gi_row - nimg * g.PH;
          const int fo = g.oy0 + irow * g.osy + nt;   // output row (STFT frame) of this tile
          asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");       // the previous tile's readers of s_w / the staging tile are done
          for (int i = et; i < 12 * 32; i += EW * 32) {
            const int kk = i / 12, gi = i - kk * 12;
            s_w[i] = gi < 9 ? __ldg(a.gain_loc + ((int64_t)nimg * 9 + gi) * 33 + kk) : 0.f;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
          float acc[4][9];
#pragma unroll
          for (int pp = 0; pp < 4; ++pp)
#pragma unroll
            for (int gi = 0; gi < 9; ++gi) acc[pp][gi] = 0.f;
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const int cc = (pp * 2 + sub) * 32 + half * 16;
              float v[16];
              tmem_ld16(tmem_row + (uint32_t)cc, v);
              if (pp == 3 && sub == 1) {               // last read of this accumulator: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(bar_tempty + 8 * b, 0); else mbar_arrive(bar_tempty + 8 * b); }
              }
              const float4* b4p = reinterpret_cast<const float4*>(a.bias + n_base + cc);
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 b4 = __ldg(b4p + (e >> 2));
                v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
              }
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float sg = sigmoid_sfu(v[e]);                         // tf.sigmoid (model.py:334)
                const float4* w4 = reinterpret_cast<const float4*>(s_w + (sub * 16 + e) * 12);
                const float4 wa = w4[0], wb = w4[1], wc = w4[2];
                const float w[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
#pragma unroll
                for (int gi = 0; gi < 9; ++gi) acc[pp][gi] = fmaf(w[gi], sg, acc[pp][gi]);
              }
            }
          }
          // write-out: per gain plane the tile is one run of 128 cells x 8 sub-pixels = 4 KB; four planes per round through the
          // staging tile, then one bulk store each
          const bool valid_f = rows_valid > 0 && (unsigned)fo < (unsigned)a.oh_lim;
#pragma unroll
          for (int r0 = 0; r0 < 9; r0 += 4) {
            if (et == 0) bulk_wait_read0();
            asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
#pragma unroll
            for (int gi = 0; gi < 4; ++gi)
              if (r0 + gi < 9)
                st_shared_v4(stile + gi * 4096u + (uint32_t)row * 32u + (uint32_t)half * 16u,
                             make_uint4(__float_as_uint(acc[0][r0 + gi]), __float_as_uint(acc[1][r0 + gi]), __float_as_uint(acc[2][r0 + gi]),
                                        __float_as_uint(acc[3][r0 + gi])));
            fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
            if (et == 0 && valid_f) {
#pragma unroll
              for (int gi = 0; gi < 4; ++gi)
                if (r0 + gi < 9)
                  bulk_s2g(a.gains + ((int64_t)nimg * 9 + r0 + gi) * a.gain_plane + (int64_t)fo * (UM_BM * 8), stile + gi * 4096u, 4096u);
              bulk_commit();
            }
          }
          continue;
        }
      }
      float4 pf[2][CPT / 4];                       // stream-K fix-up: slab values of the next pass (first two earlier pieces)
      auto slab_of = [&](int c) { return reinterpret_cast<const float4*>(a.sk_slab) + (size_t)(c * CL + (int)crank) * (UM_BM * BN / 4); };
      auto fetch_pf = [&](int c0) {
        const int np = (int)sk_c - wi.fix_first;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          if (p < np) {
            const float4* ps = slab_of(wi.fix_first + p) + (size_t)(c0 >> 5) * (CPT / 4) * (EW * 32) + et;
#pragma unroll
            for (int e4 = 0; e4 < CPT / 4; ++e4) pf[p][e4] = __ldcg(ps + (size_t)e4 * (EW * 32));
          }
        }
      };
      if (wi.fix_first >= 0) fetch_pf(0);
#pragma unroll(BN == 64 ? 2 : 1)      // (64-wide: both passes unrolled, the deferred statistics stay in registers)
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[CPT];
        long long tp0 = a.trace ? clock64() : 0;
        const int cc = c0 + half * CPT;            // first column of this thread's share of the pass
#pragma unroll
        for (int e = 0; e < CPT; e += 16) tmem_ld16(tmem_row + (uint32_t)(cc + e), v + e);
        if (CONCAT) {                              // second accumulator block: A_hi x B_lo
          float u[CPT];
#pragma unroll
          for (int e = 0; e < CPT; e += 16) tmem_ld16(tmem_row + (uint32_t)(BN + cc + e), u + e);
#pragma unroll
          for (int e = 0; e < CPT; ++e) v[e] += u[e];
        }
        if (c0 + 32 >= BN) {                       // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(bar_tempty + 8 * b, 0); else mbar_arrive(bar_tempty + 8 * b); }
        }
        if (wi.fix_first >= 0) {
          // the earlier pieces of the tile: the first two slabs were fetched a pass ahead (pf), further ones (a tile spread over
          // more than three clusters) are read here; added in cluster order
          const int np = (int)sk_c - wi.fix_first;
#pragma unroll
          for (int e4 = 0; e4 < CPT / 4; ++e4) { v[4 * e4] += pf[0][e4].x; v[4 * e4 + 1] += pf[0][e4].y; v[4 * e4 + 2] += pf[0][e4].z; v[4 * e4 + 3] += pf[0][e4].w; }
          if (np > 1) {
#pragma unroll
            for (int e4 = 0; e4 < CPT / 4; ++e4) { v[4 * e4] += pf[1][e4].x; v[4 * e4 + 1] += pf[1][e4].y; v[4 * e4 + 2] += pf[1][e4].z; v[4 * e4 + 3] += pf[1][e4].w; }
          }
          for (int c = wi.fix_first + 2; c < (int)sk_c; ++c) {
            const float4* ps = slab_of(c) + (size_t)(c0 >> 5) * (CPT / 4) * (EW * 32) + et;
            float4 f[CPT / 4];
#pragma unroll
            for (int e4 = 0; e4 < CPT / 4; ++e4) f[e4] = __ldcg(ps + (size_t)e4 * (EW * 32));
#pragma unroll
            for (int e4 = 0; e4 < CPT / 4; ++e4) { v[4 * e4] += f[e4].x; v[4 * e4 + 1] += f[e4].y; v[4 * e4 + 2] += f[e4].z; v[4 * e4 + 3] += f[e4].w; }
          }
          if (c0 + 32 < BN) fetch_pf(c0 + 32);       // the next pass's slab values travel while this pass is stored
        }
        if (!split && (a.bias != nullptr || a.relu)) {
          const int n0 = n_base + cc;
          if (a.bias != nullptr) {
            if (n0 + CPT <= a.Ntot && (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) {
#pragma unroll
              for (int e = 0; e < CPT; e += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + e));
                v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int e = 0; e < CPT; ++e)
                if (n0 + e < a.Ntot) v[e] += __ldg(a.bias + n0 + e);
            }
          }
          if (a.relu) {
#pragma unroll
            for (int e = 0; e < CPT; ++e) v[e] = fmaxf(v[e], 0.f);
          }
        }
        if (a.out_mode != 0) {
          // ---- store through the TMA unit: no shared -> register -> global copy-out.  The staging tile is written in the layout
          // the store reads (mode 1: 128-byte rows, SWIZZLE_128B like the tensor map; mode 2: four [128 rows][8 floats] blocks),
          // one elected thread issues the store(s); batch-norm column sums come from the registers (warp butterfly).
          bool st_pass = stats && !split && !(a.dbg & 1);
          const int par = (c0 >> 5) & 1;
          if (BN == 64 && defer_stats) {
            if (st_pass && row < rows_valid) {
#pragma unroll
              for (int p = 0; p < DEFER_P; ++p)
                if (p == (c0 >> 5)) {
#pragma unroll
                  for (int e = 0; e < CPT; ++e) { racc[p][e] += v[e]; rsq[p][e] = fmaf(v[e], v[e], rsq[p][e]); }
                }
            }
            st_pass = false;
          }
          if (st_pass) {
            float sq[CPT];
#pragma unroll
            for (int e = 0; e < CPT; ++e) { if (row >= rows_valid) v[e] = 0.f; sq[e] = v[e] * v[e]; }
            float sv[CPT];
#pragma unroll
            for (int e = 0; e < CPT; ++e) sv[e] = v[e];
            const float cs = warp_column_sums<CPT>(sv, lane), cq = warp_column_sums<CPT>(sq, lane);
            // s_part[par][q-th contributor of the column's half][column of the pass]: the fold below adds the four lane
            // quadrants of a column in a fixed order
            if (CPT == 32) { s_part[par][q][lane] = cs; s_part[par][q][32 + lane] = cq; }
            else if ((lane & 1) == 0) { s_part[par][q][half * 16 + (lane >> 1)] = cs; s_part[par][q][32 + half * 16 + (lane >> 1)] = cq; }
          }
          if (et == 0) bulk_wait_read0();            // the previous pass's store has finished reading the staging tile
          asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
          if (a.dbg & 4) {
          } else if (a.out_mode == 1) {
#pragma unroll
            for (int e = 0; e < CPT; e += 4) {
              const uint32_t chunk = (uint32_t)(half * CPT + e) >> 2;
              st_shared_v4(stile + (uint32_t)row * 128u + ((chunk ^ ((uint32_t)row & 7u)) << 4),
                           make_uint4(__float_as_uint(v[e]), __float_as_uint(v[e + 1]), __float_as_uint(v[e + 2]), __float_as_uint(v[e + 3])));
            }
          } else {
#pragma unroll
            for (int e = 0; e < CPT; e += 4) {
              const uint32_t col = (uint32_t)(half * CPT + e);     // column of the pass: group col / 8, float col % 8 of the row's run
              st_shared_v4(stile + (col >> 3) * 4096u + (uint32_t)row * 32u + (col & 7u) * 4u,
                           make_uint4(__float_as_uint(v[e]), __float_as_uint(v[e + 1]), __float_as_uint(v[e + 2]), __float_as_uint(v[e + 3])));
            }
          }
          fence_proxy_async();                         // generic-proxy writes -> visible to the TMA unit
          long long tp1 = a.trace ? clock64() : 0;
          asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
          if (et == 0 && rows_valid > 0 && !(a.dbg & 6)) {            // (rows_valid == 0: the all-padding tile of an odd CTA pair stores nothing)
            if (a.out_mode == 1) {
              // rows / columns beyond the tensor are clipped by the tensor map (split-K partial slabs are padded to whole tiles)
              tma_store_2d(&tm.o, stile, n_base + c0, (int)(split ? (int64_t)z * a.m_pad + m0 : m0));
            } else {
              const int gi_row = (int)(m0 / UM_BM);                   // tile = grid row (n, i): PW == 128
              const int nimg = gi_row / g.PH, irow = gi_row - nimg * g.PH;
              const int oy = g.oy0 + irow * g.osy;
              float* ybase = yf + (int64_t)nimg * g.y_sn + (int64_t)oy * g.y_sh;
#pragma unroll
              for (int gi = 0; gi < 4; ++gi) {
                const int ncol = n_base + c0 + 8 * gi;
                if (ncol < a.Ntot && (unsigned)(oy + __ldg(a.col_dy + ncol)) < (unsigned)a.oh_lim)
                  bulk_s2g(ybase + __ldg(a.col_off + ncol), stile + gi * 4096u, 4096u);
              }
            }
            bulk_commit();
          }
          if (st_pass && et < 32 && n_base + c0 + et < a.Ntot) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p) { ts += s_part[par][p][et]; tq += s_part[par][p][32 + et]; }
            s_sum[n_base + c0 + et] += ts;
            s_sqs[n_base + c0 + et] += tq;
          }
          if (a.trace) { tr_p[0] += tp1 - tp0; tr_p[3] += clock64() - tp1; tr_p[4] += 1; }
          continue;
        }
#pragma unroll
        for (int e = 0; e < CPT; e += 4)
          st_shared_v4(stile + (uint32_t)row * PITCH + (uint32_t)(half * CPT + e) * 4u,
                       make_uint4(__float_as_uint(v[e]), __float_as_uint(v[e + 1]), __float_as_uint(v[e + 2]), __float_as_uint(v[e + 3])));
        long long tp1 = a.trace ? clock64() : 0;
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (a.trace) tr_p[0] += tp1 - tp0;
        emit_pass(c0, split);
      }
    }
    if (BN == 64 && defer_stats) {
#pragma unroll
      for (int p = 0; p < DEFER_P; ++p) {
        const float cs = warp_column_sums<CPT>(racc[p], lane), cq = warp_column_sums<CPT>(rsq[p], lane);
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (CPT == 32) { s_part[0][q][lane] = cs; s_part[0][q][32 + lane] = cq; }
        else if ((lane & 1) == 0) { s_part[0][q][half * 16 + (lane >> 1)] = cs; s_part[0][q][32 + half * 16 + (lane >> 1)] = cq; }
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (et < 32 && 32 * p + et < a.Ntot) {
          float ts = 0.f, tq = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) { ts += s_part[0][k][et]; tq += s_part[0][k][32 + et]; }
          s_sum[32 * p + et] += ts;
          s_sqs[32 * p + et] += tq;
        }
      }
    }
    if (a.out_mode != 0 && et == 0) bulk_wait0();      // every store of this CTA has been written out before the CTA exits
    if (a.trace && et == 0) {
      for (int i = 0; i < 5; ++i) a.trace[blockIdx.x * 16 + 8 + i] = tr_p[i];
      { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.trace[blockIdx.x * 16 + 15] = (long long)t; }   // epilogue role done (ns)
      a.trace[blockIdx.x * 16 + 4] = tr_wait;
      a.trace[blockIdx.x * 16 + 5] = clock64() - tr_t0;
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();        // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (stats) {
    // per-CTA sums (folded above in a fixed order) -> global, as exact fixed-point integer atomics: order-independent, so the
    // statistics -- and everything downstream -- are bit-reproducible from run to run
    for (int i = tid; i < a.Ntot; i += NTHREADS) {
      fx_atomic_add(a.stat_sum + 2 * i, s_sum[i]);
      fx_atomic_add(a.stat_sqs + 2 * i, s_sqs[i]);
    }
  }
  if (a.trace && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.trace[blockIdx.x * 16 + 14] = (long long)t;            // (trace: kernel exit of this CTA, ns)
  }
}

// ---- halo-resident 3x3 convolution ----------------------------------------------------------------------------------
// The 3x3 / stride-1 / SAME convolutions of the ResNet trunk (resnet.py:224-236) through im2col copies fetch every activation
// nine times through L2 and the TMA unit (lts__t_bytes 6x the algorithmic bytes, profiles/r1_dominant_kernel_traffic.json), and
// the bytes a CTA can keep in flight in its operand ring over the L2 latency bound the whole kernel.  Here an output tile is a
// TW x TH patch of one image (TW * TH = 128 rows of the MMA, TW a multiple of 8) and, per 64-channel chunk, the producer
// fetches three TILED boxes of TW x (TH + 2) pixels -- one per horizontal tap dx, already shifted by dx, halo rows included,
// zero-filled outside the image -- each serving the three vertical taps: tap (dy, dx) is the block of 128 rows that starts
// (dy + 1) * TW rows into box dx, a 1024-byte-aligned offset, so the shared-memory descriptors are the ordinary SWIZZLE_128B
// K-major ones.  Activation bytes per tile and chunk: 3 * (TH + 2) / TH boxes instead of 9 (2.7-3.4x fewer).  Weights stream
// through their own ring (one packed chunk per tap, as in the im2col kernel).  The accumulator tile leaves through a 4-D TMA
// tensor store (box = 32 channels x TW x TH, clipped at the image border); batch-norm sums come from the registers.
// bf16x3 only (two MMAs per K step, CONCAT scheme), no bias (ResNet convolutions have none), fp32 NHWC output.
struct HaloArgs {
  const uint8_t* wpacked;
  unsigned long long* stat_sum;
  unsigned long long* stat_sqs;
  int NIMG, H, W, TW, TH, TX, TY;     // images, image size, tile shape, tiles per image
  int CC;                             // 64-channel chunks of the input
  int NDX, NDY, DX0, DY0;             // taps: dx in [DX0, DX0 + NDX) (one box each), dy in [DY0, DY0 + NDY) (blocks of a box)
  int Ntot, NT;                       // output channels, N tiles
  int SA, SB;                         // activation boxes / weight chunks in flight
  // BRES: the packed weights of all taps stay in shared memory (SB = taps * CC slots, filled by the CTA's first tile, never
  // released): layers with one N tile and few taps (conv1: 4 x 16 KB) stop re-streaming them per tile.  A1: the activation has one
  // exact bf16 plane (uint8 frames as 2k - 255, resnet18_tower): no lo box, no A_lo x B_hi MMA.
  int BRES, A1;
  long long* trace;                   // SAG_HALO_TRACE (debug): per-CTA cycle counters of the three roles
  int fast_taps;                      // resident weights: all vertical taps of a box issued by one elected block (SAG_HALO_FAST_TAPS, default 1)
  int dbg;                            // SAG_HALO_DEBUG (timing experiments, results invalid): 1 no TMEM loads in the epilogue, 2 no staging / stores
};
struct alignas(64) HaloMaps { CUtensorMap hi, lo, o, w; };     // w: tiled map of the packed weights (pairs)
constexpr int HL_THREADS = 12 * 32;    // warp 0 producer, warp 1 MMA + TMEM, warps 4-11 epilogue (two per TMEM lane quadrant)
constexpr int HL_MAX_SA = 4, HL_MAX_SB = 8;

__device__ __forceinline__ void tma_tile_4d(uint32_t dst_smem, const CUtensorMap* tmap, int c, int x, int y, int n, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(x), "r"(y), "r"(n)
               : "memory");
}
__device__ __forceinline__ void tma_tile_4d_2sm(uint32_t dst_smem, const CUtensorMap* tmap, int c, int x, int y, int n, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c), "r"(x), "r"(y), "r"(n)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, uint32_t src_smem, int c, int x, int y, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(src_smem), "r"(c), "r"(x), "r"(y), "r"(n)
               : "memory");
}

// PAIR: two CTAs of a cluster (cta_group::2) take two neighbouring tiles of the same n tile; one MMA instruction of the leader
// spans both (M = 256), each CTA feeds half of the rows of every weight operand -- rank 0 the B_hi, rank 1 the B_lo block of
// the 2*BN-wide MMA, and each its half of B_hi for the BN-wide one -- so the weight bytes per CTA drop 16 -> 12 KB per tap
// (BN = 64) and the number of MMA instructions per tile halves.  Copies of both CTAs complete on the leader's barriers.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(SAG_HL_LB, 1) halo_conv_umma_kernel(const HaloArgs a, const __grid_constant__ HaloMaps tm) {
  static_assert(BN == 64 || BN == 128, "halo kernel: 64- or 128-wide tiles (two accumulators of 2*BN TMEM columns)");
  constexpr int BX_ROWS = BN, BY_ROWS = PAIR ? BN / 2 : BN;   // pair: block X = B_hi | B_lo by rank, block Y = my half of B_hi
  constexpr int B_BYTES = (BX_ROWS + BY_ROWS) * 128;           // alone: [B_hi | B_lo] of one K chunk
  constexpr uint32_t TMEM_COLS = 4 * BN;                      // two accumulators x (hi.hi | hi.lo) blocks
  constexpr uint32_t UM_M = PAIR ? 256u : 128u;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((UM_M >> 4) << 24);
  constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((UM_M >> 4) << 24);
  constexpr int CL = PAIR ? 2 : 1;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  constexpr uint64_t DESC_HI = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);   // LBO, SBO = 1024 B, version, SWIZZLE_128B
  extern __shared__ __align__(16) uint8_t hl_smem[];
  __shared__ float s_sum[512], s_sqs[512];
  __shared__ float s_part[2][4][64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bars = smem_u32(hl_smem);
  const uint32_t bar_afull = bars, bar_aempty = bars + 8 * HL_MAX_SA;
  const uint32_t bar_bfull = bars + 16 * HL_MAX_SA, bar_bempty = bar_bfull + 8 * HL_MAX_SB;
  const uint32_t bar_tfull = bar_bempty + 8 * HL_MAX_SB, bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  const uint32_t stile0 = (bars + 256 + 1023u) & ~1023u;      // two 16 KB staging tiles (swizzled TMA stores, alternating passes)
  const uint32_t a_plane = (uint32_t)(a.TH + a.NDY - 1) * (uint32_t)a.TW * 128u;   // one bf16 plane of one box
  const uint32_t a_slot = (a.A1 ? 1u : 2u) * a_plane;
  const uint32_t a_base = stile0 + 32768u;
  const uint32_t b_base = a_base + (uint32_t)a.SA * a_slot;
  const int SA = a.SA, SB = a.SB;
  const int tiles = a.NIMG * a.TY * a.TX;
  const int tile_groups = (tiles + CL - 1) / CL;             // pair: two neighbouring tiles per cluster (the last may be padding)
  const int n_work = tile_groups * a.NT;                      // work items per cluster
  const int wk0 = blockIdx.x / CL, wk_step = gridDim.x / CL;

  for (int i = tid; i < 512; i += HL_THREADS) { s_sum[i] = 0.f; s_sqs[i] = 0.f; }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < SA; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
      for (int s = 0; s < SB; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 8 * CL); }
      fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) tmem_alloc2(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  pdl_prologue();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // work item -> (tile of one image, n tile): tiles fastest, so neighbouring CTAs share the weight chunks in L2
  auto decode = [&](int wk, int& nimg, int& y0, int& x0, int& nt, bool& real) {
    int t = (wk % tile_groups) * CL + (int)crank;
    nt = wk / tile_groups;
    real = t < tiles;                                         // (the padding tile of an odd pair reloads the last tile, stores nothing)
    if (!real) t = tiles - 1;
    const int per = a.TY * a.TX;
    nimg = t / per;
    const int r = t - nimg * per;
    y0 = (r / a.TX) * a.TH;
    x0 = (r % a.TX) * a.TW;
  };

  if (warp == 0) {
    // ================================ producer ================================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t lead_bars = PAIR ? map_to_cta(bars, 0) : bars;             // the leader's barriers (cluster address)
    const uint32_t lead_afull = lead_bars + (bar_afull - bars), lead_bfull = lead_bars + (bar_bfull - bars);
    long long tr_wait = 0;
    const long long tr_t0 = clock64();
    for (int wk = wk0; wk < n_work; wk += wk_step) {
      int nimg, y0, x0, nt;
      bool real;
      decode(wk, nimg, y0, x0, nt, real);
      const bool load_b = !a.BRES || wk == wk0;             // resident weights: the first tile fills the slots
      for (int c = 0; c < a.CC; ++c) {
#pragma unroll 1
        for (int dx = 0; dx < a.NDX; ++dx) {
          const long long tw0 = a.trace ? clock64() : 0;
          mbar_wait(bar_aempty + 8 * sa, pa ^ 1);
          if (a.trace) tr_wait += clock64() - tw0;
          if (elect_one()) {
            const uint32_t dst = a_base + (uint32_t)sa * a_slot;
            if (!PAIR) {
              const uint32_t bar = bar_afull + 8 * sa;
              mbar_arrive_expect_tx(bar, a_slot);
              tma_tile_4d(dst, &tm.hi, c * 64, x0 + dx + a.DX0, y0 + a.DY0, nimg, bar);
              if (!a.A1) tma_tile_4d(dst + a_plane, &tm.lo, c * 64, x0 + dx + a.DX0, y0 + a.DY0, nimg, bar);
            } else {
              // the leader's producer makes the single arrival and expects the bytes of both CTAs (see the im2col kernel)
              const uint32_t bar = lead_afull + 8 * sa;
              if (crank == 0) mbar_arrive_expect_tx(bar_afull + 8 * sa, 2 * a_slot);
              tma_tile_4d_2sm(dst, &tm.hi, c * 64, x0 + dx + a.DX0, y0 + a.DY0, nimg, bar);
              if (!a.A1) tma_tile_4d_2sm(dst + a_plane, &tm.lo, c * 64, x0 + dx + a.DX0, y0 + a.DY0, nimg, bar);
            }
          }
          __syncwarp();
          if (++sa == SA) { sa = 0; pa ^= 1; }
#pragma unroll 1
          for (int dy = 0; dy < a.NDY; ++dy) {
            if (!load_b) continue;
            mbar_wait(bar_bempty + 8 * sb, pb ^ 1);
            if (elect_one()) {
              const int kc = (dy * a.NDX + dx) * a.CC + c;     // packed K order: tap-major (rows, then columns), then channel chunk
              const uint32_t dst = b_base + (uint32_t)sb * B_BYTES;
              if (!PAIR) {
                const uint32_t bar = bar_bfull + 8 * sb;
                mbar_arrive_expect_tx(bar, B_BYTES);
                bulk_g2s(dst, a.wpacked + ((size_t)nt * (a.NDX * a.NDY * a.CC) + kc) * (size_t)(2 * BN * 128), B_BYTES, bar);
              } else {
                const uint32_t bar = lead_bfull + 8 * sb;
                if (crank == 0) mbar_arrive_expect_tx(bar_bfull + 8 * sb, 2 * B_BYTES);
                const int wrow = (nt * (a.NDX * a.NDY * a.CC) + kc) * (2 * BN);     // rows of 128 bytes; boxes of BN/2 rows
                constexpr int HB = BN / 2;
                tma_tile_2d_2sm(dst, &tm.w, 0, wrow + (int)crank * BN, bar);                       // block X: B_hi (rank 0) / B_lo (rank 1)
                tma_tile_2d_2sm(dst + HB * 128, &tm.w, 0, wrow + (int)crank * BN + HB, bar);
                tma_tile_2d_2sm(dst + BX_ROWS * 128, &tm.w, 0, wrow + (int)crank * HB, bar);      // block Y: my half of B_hi
              }
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
    if (a.trace && lane == 0) { a.trace[blockIdx.x * 8 + 0] = clock64() - tr_t0; a.trace[blockIdx.x * 8 + 1] = tr_wait; }
  } else if (warp == 1) {
    // ================================ MMA issuer (pair: the leader only) ================================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it_local = 0;
    long long tr_wa = 0, tr_wt = 0;
    const long long tr_t0 = clock64();
    if (!PAIR || crank == 0)
    for (int wk = wk0; wk < n_work; wk += wk_step, ++it_local) {
      const int b = it_local & 1;
      const uint32_t use = (uint32_t)(it_local >> 1);
      const long long tw0 = a.trace ? clock64() : 0;
      mbar_wait(bar_tempty + 8 * b, (use & 1) ^ 1);
      if (a.trace) tr_wt += clock64() - tw0;
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(b * 2 * BN);
      uint32_t first = 0;
      const bool wait_b = !a.BRES || it_local == 0;         // resident weights: slot sb = tap, filled once
      for (int c = 0; c < a.CC; ++c) {
#pragma unroll 1
        for (int dx = 0; dx < a.NDX; ++dx) {
          const long long tw1 = a.trace ? clock64() : 0;
          mbar_wait(bar_afull + 8 * sa, pa);
          if (a.trace) tr_wa += clock64() - tw1;
          tc_fence_after();
          const uint32_t slot = a_base + (uint32_t)sa * a_slot;
          // Resident weights (conv1): nothing to wait for between the vertical taps, so ONE elected block issues their NDY x 4 (x 2)
          // MMAs back to back with descriptors advanced by addition.  The general loop below spends ~870 clk of scalar work per
          // tap (elect, descriptor assembly, commits, warp sync) -- more than the 256 clk its four N = 128 MMAs take.
          if (a.fast_taps && !wait_b && sb + a.NDY <= SB) {
            if (elect_one()) {
              uint64_t da_hi = DESC_HI | (uint64_t)((slot & 0x3FFFFu) >> 4);
              uint64_t da_lo = DESC_HI | (uint64_t)(((slot + a_plane) & 0x3FFFFu) >> 4);
              uint64_t db = DESC_HI | (uint64_t)(((b_base + (uint32_t)sb * B_BYTES) & 0x3FFFFu) >> 4);
              uint64_t dby = DESC_HI | (uint64_t)(((b_base + (uint32_t)sb * B_BYTES + BX_ROWS * 128) & 0x3FFFFu) >> 4);
              const uint64_t step_a = (uint64_t)(((uint32_t)a.TW * 128u) >> 4), step_b = (uint64_t)(B_BYTES >> 4);
#pragma unroll 1
              for (int dy = 0; dy < a.NDY; ++dy) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  if (PAIR) {
                    umma_bf16_2(tmem_acc, da_hi + 2 * k4, db + 2 * k4, IDESC2, first | (uint32_t)(k4 > 0));
                    if (!a.A1) umma_bf16_2(tmem_acc, da_lo + 2 * k4, dby + 2 * k4, IDESC, 1u);
                  } else {
                    umma_bf16(tmem_acc, da_hi + 2 * k4, db + 2 * k4, IDESC2, first | (uint32_t)(k4 > 0));
                    if (!a.A1) umma_bf16(tmem_acc, da_lo + 2 * k4, db + 2 * k4, IDESC, 1u);
                  }
                }
                first = 1u;
                da_hi += step_a; da_lo += step_a; db += step_b; dby += step_b;
              }
              const bool last = dx + 1 == a.NDX && c + 1 == a.CC;
              if (PAIR) {
                umma_commit2(bar_aempty + 8 * sa);
                if (last) umma_commit2(bar_tfull + 8 * b);
              } else {
                umma_commit(bar_aempty + 8 * sa);
                if (last) umma_commit(bar_tfull + 8 * b);
              }
            }
            __syncwarp();
            first = 1u;
            sb += a.NDY;
            if (sb >= SB) { sb -= SB; pb ^= 1; }
            if (++sa == SA) { sa = 0; pa ^= 1; }
            continue;
          }
#pragma unroll 1
          for (int dy = 0; dy < a.NDY; ++dy) {
            if (wait_b) {                                  // (resident weights: no wait, and no fence between the taps' MMAs)
              mbar_wait(bar_bfull + 8 * sb, pb);
              tc_fence_after();
            }
            const uint32_t ab = slot + (uint32_t)dy * (uint32_t)a.TW * 128u;        // rows dy * TW .. of the box: vertical tap dy - 1
            const uint32_t bb = b_base + (uint32_t)sb * B_BYTES;
            const uint64_t da_hi = DESC_HI | (uint64_t)((ab & 0x3FFFFu) >> 4);
            const uint64_t da_lo = DESC_HI | (uint64_t)(((ab + a_plane) & 0x3FFFFu) >> 4);
            const uint64_t db = DESC_HI | (uint64_t)((bb & 0x3FFFFu) >> 4);
            const uint64_t dby = DESC_HI | (uint64_t)(((bb + BX_ROWS * 128) & 0x3FFFFu) >> 4);    // pair: the halves of B_hi
            if (elect_one()) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                if (PAIR) {
                  umma_bf16_2(tmem_acc, da_hi + 2 * k4, db + 2 * k4, IDESC2, first | (uint32_t)(k4 > 0));
                  if (!a.A1) umma_bf16_2(tmem_acc, da_lo + 2 * k4, dby + 2 * k4, IDESC, 1u);
                } else {
                  umma_bf16(tmem_acc, da_hi + 2 * k4, db + 2 * k4, IDESC2, first | (uint32_t)(k4 > 0));   // [A_hi.B_hi | A_hi.B_lo]
                  if (!a.A1) umma_bf16(tmem_acc, da_lo + 2 * k4, db + 2 * k4, IDESC, 1u);                 // += A_lo.B_hi
                }
              }
              const bool last = dy + 1 == a.NDY && dx + 1 == a.NDX && c + 1 == a.CC;
              if (PAIR) {
                if (!a.BRES) umma_commit2(bar_bempty + 8 * sb);
                if (dy + 1 == a.NDY) umma_commit2(bar_aempty + 8 * sa);
                if (last) umma_commit2(bar_tfull + 8 * b);
              } else {
                if (!a.BRES) umma_commit(bar_bempty + 8 * sb);
                if (dy + 1 == a.NDY) umma_commit(bar_aempty + 8 * sa);
                if (last) umma_commit(bar_tfull + 8 * b);
              }
            }
            __syncwarp();
            first = 1u;
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
      }
    }
    if (a.trace && lane == 0 && (!PAIR || crank == 0)) {
      a.trace[blockIdx.x * 8 + 2] = clock64() - tr_t0; a.trace[blockIdx.x * 8 + 3] = tr_wa; a.trace[blockIdx.x * 8 + 4] = tr_wt;
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    long long tr_we = 0;
    const long long tr_t0 = clock64();
    const int q = warp & 3, half = warp >= 8 ? 0 : 1;     // TMEM lane quadrant = warp % 4; two warps per quadrant, 16 of a pass's 32 columns each
    const int et = (half * 4 + q) * 32 + lane;
    const int row = q * 32 + lane;
    const int rr = row / a.TW, rc = row - rr * a.TW;      // position of this row's pixel inside the tile
    int it_local = 0;
    uint32_t pass_no = 0;                                 // passes of this CTA so far: alternates the staging tile
    // 64-wide tiles, one N tile (conv2_x): the batch-norm sums of a thread's 2 x 16 columns stay in registers over all tiles of
    // the CTA and cross the warp once, after the last tile
    constexpr int DEFER_P = BN == 64 ? 2 : 1;
    const bool defer_stats = BN == 64 && a.NT == 1;
    float racc[DEFER_P][16], rsq[DEFER_P][16];
#pragma unroll
    for (int p = 0; p < DEFER_P; ++p)
#pragma unroll
      for (int e = 0; e < 16; ++e) { racc[p][e] = 0.f; rsq[p][e] = 0.f; }
    for (int wk = wk0; wk < n_work; wk += wk_step, ++it_local) {
      int nimg, y0, x0, nt;
      bool real;
      decode(wk, nimg, y0, x0, nt, real);
      const int b = it_local & 1;
      const uint32_t use = (uint32_t)(it_local >> 1);
      const bool valid = real && y0 + rr < a.H && x0 + rc < a.W;   // pixels of a border tile beyond the image: clipped by the store, masked here
      const int n_base = nt * BN;
      const long long tw0 = a.trace ? clock64() : 0;
      mbar_wait(bar_tfull + 8 * b, use & 1);
      if (a.trace) tr_we += clock64() - tw0;
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + (uint32_t)(b * 2 * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll(BN == 64 ? 2 : 1)      // (64-wide: both passes unrolled, the deferred statistics stay in registers)
      for (int c0 = 0; c0 < BN; c0 += 32, ++pass_no) {
        const int cc = c0 + half * 16;
        const int par = (int)(pass_no & 1u);
        const uint32_t stile = stile0 + (uint32_t)par * 16384u;
        float v[16], u[16];
        if (a.dbg & 1) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { v[e] = 0.f; u[e] = 0.f; }
        } else {
          tmem_ld16_nowait(tmem_row + (uint32_t)cc, v);              // both accumulator blocks in flight, one wait
          tmem_ld16_nowait(tmem_row + (uint32_t)(BN + cc), u);
          tmem_ld_wait();
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] += u[e];
        if (c0 + 32 >= BN) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(bar_tempty + 8 * b, 0); else mbar_arrive(bar_tempty + 8 * b); }
        }
        if (BN == 64 && defer_stats) {
          if (valid) {
#pragma unroll
            for (int p = 0; p < DEFER_P; ++p)
              if (p == (c0 >> 5)) {
#pragma unroll
                for (int e = 0; e < 16; ++e) { racc[p][e] += v[e]; rsq[p][e] = fmaf(v[e], v[e], rsq[p][e]); }
              }
          }
        } else {
          float sv[16], sq[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { sv[e] = valid ? v[e] : 0.f; sq[e] = sv[e] * sv[e]; }
          const float cs = warp_column_sums<16>(sv, lane), cq = warp_column_sums<16>(sq, lane);
          if ((lane & 1) == 0) { s_part[par][q][half * 16 + (lane >> 1)] = cs; s_part[par][q][32 + half * 16 + (lane >> 1)] = cq; }
        }
        // ONE barrier per pass, two staging tiles: the store that last read THIS tile (two passes ago) was waited for by the
        // issuing thread before it arrived at the previous pass's barrier (wait_group.read 0 below: at that point the youngest
        // store is a whole pass old, so the wait is rarely a stall and the store latency does not sit between two passes)
        if (!(a.dbg & 2)) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const uint32_t chunk = (uint32_t)(half * 16 + e) >> 2;
            st_shared_v4(stile + (uint32_t)row * 128u + ((chunk ^ ((uint32_t)row & 7u)) << 4),
                         make_uint4(__float_as_uint(v[e]), __float_as_uint(v[e + 1]), __float_as_uint(v[e + 2]), __float_as_uint(v[e + 3])));
          }
        }
        fence_proxy_async();
        if (et == 0) bulk_wait_read0();                    // the previous pass's store has read the OTHER tile: the next pass may write it
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          if (real && !(a.dbg & 2)) tma_store_4d(&tm.o, stile, n_base + c0, x0, y0, nimg);
          bulk_commit();                                   // (one group per pass, also an empty one: wait_group counts groups)
        }
        if (!(BN == 64 && defer_stats) && et < 32 && n_base + c0 + et < a.Ntot) {
          float ts = 0.f, tq = 0.f;
#pragma unroll
          for (int p = 0; p < 4; ++p) { ts += s_part[par][p][et]; tq += s_part[par][p][32 + et]; }
          s_sum[n_base + c0 + et] += ts;
          s_sqs[n_base + c0 + et] += tq;
        }
      }
    }
    if (BN == 64 && defer_stats) {
#pragma unroll
      for (int p = 0; p < DEFER_P; ++p) {
        const float cs = warp_column_sums<16>(racc[p], lane), cq = warp_column_sums<16>(rsq[p], lane);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if ((lane & 1) == 0) { s_part[0][q][half * 16 + (lane >> 1)] = cs; s_part[0][q][32 + half * 16 + (lane >> 1)] = cq; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et < 32 && 32 * p + et < a.Ntot) {
          float ts = 0.f, tq = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) { ts += s_part[0][k][et]; tq += s_part[0][k][32 + et]; }
          s_sum[32 * p + et] += ts;
          s_sqs[32 * p + et] += tq;
        }
      }
    }
    if (et == 0) bulk_wait0();
    if (a.trace && et == 0) { a.trace[blockIdx.x * 8 + 5] = clock64() - tr_t0; a.trace[blockIdx.x * 8 + 6] = tr_we; }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();          // no CTA leaves while the peer may still signal its barriers or read its operands
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (a.stat_sum != nullptr) {
    for (int i = tid; i < a.Ntot; i += HL_THREADS) {
      fx_atomic_add(a.stat_sum + 2 * i, s_sum[i]);
      fx_atomic_add(a.stat_sqs + 2 * i, s_sqs[i]);
    }
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

// GEMM column n of a sub-pixel transposed conv -> (sub-pixel row py, sub-pixel column px, output channel co)
//   order 0: n = (py*sw + px)*Cout + co                  NHWC outputs: a pixel's channels are contiguous
//   order 1: n = (py*Cout + co)*sw + px                  planar outputs: a row's sub-pixels are contiguous
//   order 2: (sw == 8, Cout == 32) n = py*256 + c, c = pass*32 + half*16 + e -> px = 4*half + pass/2, co = 16*(pass%2) + e:
//            the epilogue thread that drains columns [half*16, half*16+16) of every 32-column pass of a 256-wide tile
//            sees, over two consecutive passes, all 32 tracks of ONE sub-pixel -- and over the eight passes four
//            neighbouring sub-pixels (the mask-gain fusion of out_mode 3)
__host__ __device__ inline void decode_subpixel_column(int order, int n, int cout, int sw, int* py, int* px, int* co) {
  if (order == 0) { *co = n % cout; const int ph = n / cout; *px = ph % sw; *py = ph / sw; }
  else if (order == 1) { *px = n % sw; const int q = n / sw; *co = q % cout; *py = q / cout; }
  else { *py = n >> 8; const int c = n & 255, pass = c >> 5, half = (c >> 4) & 1, e = c & 15; *px = 4 * half + (pass >> 1); *co = ((pass & 1) << 4) + e; }
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
    decode_subpixel_column(order, n, cout, sw, &py, &px, &co);
    const int p = py + sh * ta, qq = px + sw * tb;
    wk[idx] = (p < kh && qq < kw) ? __ldg(w + (((int64_t)p * kw + qq) * cout + co) * cin + ci) : 0.f;
  }
}

// HWIO weights of a stride-2 convolution over C channels -> weights of the equivalent stride-1 convolution over the 2x2
// space-to-depth image with 16 channels per pixel (channel (py*2+px)*C + c, zero beyond 4*C):
//   out[(a*tw + b)*16 + ch][co] = w[2a+py][2b+px][c][co]   (zero when the kernel index falls outside kh x kw)
__global__ void s2d_weights_kernel(const float* __restrict__ w, int kh, int kw, int cin, int cout, int th, int tw, float div,
                                   float* __restrict__ out) {
  const int64_t total = (int64_t)th * tw * 16 * cout;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int co = (int)(idx % cout);
    int64_t r = idx / cout;
    const int ch = (int)(r % 16); r /= 16;
    const int b = (int)(r % tw);
    const int a = (int)(r / tw);
    float v = 0.f;
    if (ch < 4 * cin) {
      const int sub = ch / cin, c = ch - sub * cin;
      const int ky = 2 * a + (sub >> 1), kx = 2 * b + (sub & 1);
      if (ky < kh && kx < kw) v = __ldg(w + (((int64_t)ky * kw + kx) * cin + c) * cout + co) / div;
    }
    out[idx] = v;
  }
}

__global__ void expand_bias_kernel(const float* __restrict__ bias, int cout, int sw, int N, int order, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int py, px, co;
  decode_subpixel_column(order, n, cout, sw, &py, &px, &co);
  out[n] = __ldg(bias + co);
}

// ---- split-K finish: sum the partial accumulators, then the same bias / activation / statistics / store as the
//      fused epilogue.  A block owns SK_ROWS (1..16) consecutive rows; its 256 threads sweep (row, 4-column group) pairs with the
//      column group fastest, so partial reads and output writes are whole contiguous rows; when 256 % groups == 0 a
//      thread keeps one column group and carries that group's batch-norm sums in registers. ----
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const __grid_constant__ GatherGeom g, const UmmaArgs a, int Z,
                                                            int SK_ROWS) {
  __shared__ float s_sum[UM_MAX_N], s_sqs[UM_MAX_N];
  pdl_prologue();
  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  const int64_t r0 = (int64_t)blockIdx.x * SK_ROWS;
  const int rows = (int)((M - r0) < SK_ROWS ? (M - r0) : SK_ROWS);
  const int ncg = (a.Ntot + 3) / 4;
  const bool stats = a.stat_sum != nullptr;
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssqs[4] = {0.f, 0.f, 0.f, 0.f};
  float* yf = reinterpret_cast<float*>(a.y);
  const int total = rows * ncg;
  const int64_t zs = a.m_pad * a.n_pad;
  for (int base = threadIdx.x; base < total; base += 4 * blockDim.x) {
   // four (row, column group) items per thread per round: all partial loads are issued before any is consumed
   float4 accs[4];
#pragma unroll
   for (int u = 0; u < 4; ++u) {
     accs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
     const int idx = base + u * blockDim.x;
     if (idx < total) {
       const int rr = idx / ncg, cg = idx - rr * ncg;
       const float* pp = a.partial + (r0 + rr) * a.n_pad + cg * 4;
#pragma unroll 4
       for (int z = 0; z < Z; ++z) {
         const float4 p = __ldcs(reinterpret_cast<const float4*>(pp + (int64_t)z * zs));
         accs[u].x += p.x; accs[u].y += p.y; accs[u].z += p.z; accs[u].w += p.w;
       }
     }
   }
#pragma unroll
   for (int u = 0; u < 4; ++u) {
    const int idx = base + u * blockDim.x;
    if (idx >= total) continue;
    const int rr = idx / ncg, cg = idx - rr * ncg;
    const int n = cg * 4;
    const int64_t m = r0 + rr;
    const float4 acc = accs[u];
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (a.bias != nullptr && n + e < a.Ntot) v[e] += __ldg(a.bias + n + e);
      if (a.relu) v[e] = fmaxf(v[e], 0.f);
      if (n + e >= a.Ntot) v[e] = 0.f;
    }
    int64_t yoff;
    int oy = 0, ox = 0;
    if (a.dense) {
      yoff = m * g.y_sw;
    } else {
      const uint32_t mu = (uint32_t)m;
      const uint32_t qq = mu / (uint32_t)g.PW, j = mu - qq * (uint32_t)g.PW;
      const uint32_t b = qq / (uint32_t)g.PH, i = qq - b * (uint32_t)g.PH;
      oy = g.oy0 + (int)i * g.osy;
      ox = g.ox0 + (int)j * g.osx;
      yoff = (int64_t)b * g.y_sn + (int64_t)oy * g.y_sh + (int64_t)ox * g.y_sw;
    }
    if (a.vec_store) {
      bool ok = true;
      int64_t eoff = yoff + n;
      if (a.col_off != nullptr) {
        ok = (unsigned)(oy + a.col_dy[n]) < (unsigned)a.oh_lim && (unsigned)(ox + a.col_dx[n]) < (unsigned)a.ow_lim;
        eoff = yoff + a.col_off[n];
      }
      if (ok) {
        if (a.out_bf2) store_bf2_4(a.y, a.y_plane, eoff, v, a.out_bf2);
        else *reinterpret_cast<float4*>(yf + eoff) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n + e >= a.Ntot) continue;
        int64_t eoff = yoff + (int64_t)(n + e) * g.y_sc;
        if (a.col_off != nullptr) {
          if (!((unsigned)(oy + a.col_dy[n + e]) < (unsigned)a.oh_lim && (unsigned)(ox + a.col_dx[n + e]) < (unsigned)a.ow_lim)) continue;
          eoff = yoff + a.col_off[n + e];
        }
        if (a.out_bf2) store_bf2_1(a.y, a.y_plane, eoff, v[e], a.out_bf2);
        else yf[eoff] = v[e];
      }
    }
    if (stats) {          // (the host only asks for statistics when 256 % column groups == 0: a thread keeps one column group)
#pragma unroll
      for (int e = 0; e < 4; ++e) { ssum[e] += v[e]; ssqs[e] = fmaf(v[e], v[e], ssqs[e]); }
    }
   }
  }
  if (stats) {
    // block sums in a fixed order: every thread's register sums go to shared memory, thread c < ncg adds the 256/ncg
    // contributors of column group c in thread order; across blocks: exact fixed-point atomics (order-independent)
    __shared__ float s_tp[256][8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { s_tp[threadIdx.x][e] = ssum[e]; s_tp[threadIdx.x][4 + e] = ssqs[e]; }
    __syncthreads();
    if ((int)threadIdx.x < ncg) {
      float ts[4] = {0.f, 0.f, 0.f, 0.f}, tq[4] = {0.f, 0.f, 0.f, 0.f};
      for (int t = threadIdx.x; t < (int)blockDim.x; t += ncg) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { ts[e] += s_tp[t][e]; tq[e] += s_tp[t][4 + e]; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if ((int)threadIdx.x * 4 + e < a.Ntot) { s_sum[threadIdx.x * 4 + e] = ts[e]; s_sqs[threadIdx.x * 4 + e] = tq[e]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.Ntot; i += blockDim.x) {
      fx_atomic_add(a.stat_sum + 2 * i, s_sum[i]);
      fx_atomic_add(a.stat_sqs + 2 * i, s_sqs[i]);
    }
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int reduce_rows_per_block(int64_t M) {
  int64_t rpb = M / (2 * num_sms());              // rows per block: keep >= 2 blocks per SM, at most 16 rows
  if (rpb > 16) rpb = 16;
  if (rpb < 1) rpb = 1;
  return (int)rpb;
}
int max_conv_ctas() {
  static const int max_ctas = env_int("SAG_UMMA_MAX_CTAS", 0);
  return max_ctas > 0 ? max_ctas : num_sms();
}

template <int BN, int NSPLIT, int SRC, bool PAIR>
int launch_cfg(const GatherGeom& g, const UmmaArgs& a_in, const TmaPair& tm, int nt, int Z, cudaStream_t st) {
  constexpr int PLANES = NSPLIT >= 2 ? 2 : 1;
  constexpr int BX_ROWS = !PAIR ? BN : (NSPLIT == 2 ? BN : BN / 2);          // (same as in the kernel)
  constexpr int BY_ROWS = PLANES == 1 ? 0 : (!PAIR ? BN : BN / 2);
  constexpr int STAGE_BYTES = PLANES * UM_A_PLANE + (BX_ROWS + BY_ROWS) * 128;
  UmmaArgs a = a_in;
  a.NT = nt;
  a.Z = Z;
  // one persistent CTA per SM: barriers + staging tile + alignment slack + as many operand stages as fit (<= 6) in what the
  // static shared memory of this instantiation leaves of the per-block opt-in maximum (227 KB on sm_100)
  auto kern = gather_gemm_umma_kernel<BN, NSPLIT, SRC, PAIR>;
  static int budget[64] = {0};                      // per device: the attribute lives in the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (budget[dev & 63] == 0) {
    cudaFuncAttributes fa;
    int optin = 0;
    SAG_CHECK_CUDA(cudaFuncGetAttributes(&fa, kern));
    SAG_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int b = optin - (int)fa.sharedSizeBytes - env_int("SAG_UMMA_SMEM_RESERVE", 0);   // (development: shared memory left to co-resident CTAs of other kernels)
    SAG_REQUIRE(b > 64 * 1024, SAG_ECUDA, "tcgen05 path: only %d bytes of dynamic shared memory available", b);
    SAG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    budget[dev & 63] = b;
  }
  const int fixed = UM_BAR_BYTES + 1024 + UM_STAGING_BYTES + 1024;
  int S = (budget[dev & 63] - fixed) / STAGE_BYTES;
  if (S > 6) S = 6;    // the cp.async drain handles at most 5 groups in flight
  if (S < 2) S = 2;
  a.stages = S;
  const size_t smem = (size_t)fixed + (size_t)S * STAGE_BYTES;
  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  const int64_t MT = cdiv64(M, UM_BM);
  constexpr int CL = PAIR ? 2 : 1;                                  // CTA pair: a cluster of two SMs of one TPC per work item
  const int64_t n_work = cdiv64(MT, CL) * nt * Z;                   // per cluster
  int64_t clusters = max_conv_ctas() / CL;                          // (SAG_UMMA_MAX_CTAS: test knob, forces many work items per CTA)
  if (clusters < 1) clusters = 1;
  if (a.sk_flags != nullptr) {
    // stream-K: every cluster gets work (>= 2 units each), a CTA has one slab and one flag
    if (clusters * CL > UMMA_SK_FLAGS) clusters = UMMA_SK_FLAGS / CL;
    if (clusters * CL > num_sms()) clusters = num_sms() / CL;
    if (n_work * a.KC < 2 * clusters) clusters = std::max<int64_t>(1, n_work * a.KC / 2);
  } else if (n_work < clusters) {
    clusters = n_work;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  cfg.blockDim = dim3(um_threads(SRC));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = CL;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cudaLaunchAttribute attrs[2];
  attrs[0] = attr;
  memset(&attrs[1], 0, sizeof(attrs[1]));
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (PAIR) {
    // a persistent grid must be co-resident: ask the driver how many 2-CTA clusters of this kernel fit at once (GPCs with an
    // odd number of free SMs, SMs held by other work) and never launch more -- a cluster that has to wait for a second wave
    // doubles the kernel's time
    static int max_active[64] = {0};
    if (max_active[dev & 63] == 0) {
      int n = 0;
      cfg.gridDim = dim3((unsigned)(num_sms() / CL * CL));
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = num_sms() / CL; }
      max_active[dev & 63] = n;
      if (env_int("SAG_UMMA_DEBUG", 0)) fprintf(stderr, "[umma] BN=%d NSPLIT=%d pair: max active clusters %d (smem %zu B)\n", BN, NSPLIT, n, smem);
    }
    if (clusters > max_active[dev & 63]) clusters = max_active[dev & 63];
    cfg.gridDim = dim3((unsigned)(clusters * CL));
  }
  // debug: SAG_UMMA_TRACE=<M tiles> prints where the three roles of the first launch with that many M tiles wait
  static const int trace_mt = env_int("SAG_UMMA_TRACE", 0);
  static int traced = 0;
  if (trace_mt > 0 && MT == trace_mt && traced < env_int("SAG_UMMA_TRACE_N", 1)) {
    ++traced;
    const size_t n = (size_t)cfg.gridDim.x * 16;
    long long* dtr = nullptr;
    cudaMalloc(&dtr, n * sizeof(long long));
    cudaMemset(dtr, 0, n * sizeof(long long));
    a.trace = dtr;
    cudaStreamSynchronize(st);
    cudaLaunchKernelEx(&cfg, kern, g, a, tm);
    cudaStreamSynchronize(st);
    std::vector<long long> tr(n);
    cudaMemcpy(tr.data(), dtr, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dtr);
    double d[16] = {0};
    for (size_t c = 0; c < n / 16; ++c)
      for (int i = 0; i < 16; ++i) d[i] += (double)tr[c * 16 + i] / (double)(n / 16);
    fprintf(stderr, "[umma trace] BN=%d NSPLIT=%d SRC=%d ctas=%u MT=%lld NT=%d Z=%d KC=%d stages=%d chunks/cta=%.1f\n", BN, NSPLIT, SRC,
            cfg.gridDim.x, (long long)MT, nt, Z, a.KC, S, d[6]);
    fprintf(stderr, "[umma trace]   MMA thread   : total %9.0f clk, waiting for operands %9.0f, for a free accumulator %9.0f\n", d[1], d[0], d[7]);
    fprintf(stderr, "[umma trace]   producer t0  : total %9.0f clk, waiting for a free stage %9.0f\n", d[3], d[2]);
    fprintf(stderr, "[umma trace]   epilogue t0  : total %9.0f clk, waiting for an accumulator %9.0f\n", d[5], d[4]);
    {
      long long e0 = LLONG_MAX, e1 = 0, x0 = LLONG_MAX, x1 = 0, r0 = LLONG_MAX, r1 = 0;
      for (size_t c = 0; c < n / 16; ++c) {
        e0 = std::min(e0, tr[c * 16 + 13]); e1 = std::max(e1, tr[c * 16 + 13]);
        x0 = std::min(x0, tr[c * 16 + 14]); x1 = std::max(x1, tr[c * 16 + 14]);
        r0 = std::min(r0, tr[c * 16 + 15]); r1 = std::max(r1, tr[c * 16 + 15]);
      }
      fprintf(stderr, "[umma trace]   wall (us from the first CTA's entry): last entry %.1f, epilogue roles done %.1f .. %.1f, CTA exits %.1f .. %.1f\n",
              (e1 - e0) * 1e-3, (r0 - e0) * 1e-3, (r1 - e0) * 1e-3, (x0 - e0) * 1e-3, (x1 - e0) * 1e-3);
    }
    fprintf(stderr, "[umma trace]   epilogue t0  : tmem->smem %9.0f, copy-out %9.0f, statistics %9.0f, barriers %9.0f (sum over %0.f passes)\n",
            d[8], d[9], d[10], d[11], d[12]);
    SAG_LAUNCH_CHECK();
    return SAG_OK;
  }
  SAG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, g, a, tm));
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

template <int BN, int NSPLIT>
int launch_ns(const GatherGeom& g, const UmmaArgs& a, const TmaPair& tm, int nt, int src, int Z, cudaStream_t st) {
  switch (src) {
    case SRC_TMA:
      if constexpr (BN >= 64) {
        // CTA pairs (cta_group::2, M = 256): two adjacent M tiles per cluster, each CTA feeds half of every weight tile, so
        // the weight bytes pulled through L2 and written to / read from shared memory halve.  Both CTAs' copies complete on
        // the leader's barrier (2-SM TMA), commits are multicast.  Measured on B200 (profiles/README.md, round 2): 256-wide
        // tiles with a long K loop gain 11 % (conv4_x 55.6 -> 49.5 us); 64- / 128-wide tiles are bound by the activation
        // bytes, which pairs do not reduce (+-0 %), and short K loops lose 13-19 % to the cluster launch / lockstep.
        // Default (-1 / unset): pairs exactly where they win.  SAG_UMMA_PAIR=0/1 or sag_set_option "cta_pair" force it.
        static const int pair_env = env_int("SAG_UMMA_PAIR", -1);
        const int64_t MT = cdiv64((int64_t)g.N * g.PH * g.PW, UM_BM);
        const int want = g_umma_pair >= 0 ? g_umma_pair : pair_env;
        const bool pair = want >= 0 ? want != 0 : (BN == 256 && a.KC / Z >= 16);
        if (pair && MT >= 2 && a.pair_ok) return launch_cfg<BN, NSPLIT, SRC_TMA, true>(g, a, tm, nt, Z, st);
        return launch_cfg<BN, NSPLIT, SRC_TMA, false>(g, a, tm, nt, Z, st);
      } else {
        return launch_cfg<BN, NSPLIT, SRC_BF2, false>(g, a, tm, nt, Z, st);  // (the host never picks TMA for 32-wide tiles)
      }
    case SRC_BF2: return launch_cfg<BN, NSPLIT, SRC_BF2, false>(g, a, tm, nt, Z, st);
    case SRC_F32_VEC: return launch_cfg<BN, NSPLIT, SRC_F32_VEC, false>(g, a, tm, nt, Z, st);
    default: return launch_cfg<BN, NSPLIT, SRC_F32, false>(g, a, tm, nt, Z, st);
  }
}

template <int BN>
int launch_bn(const GatherGeom& g, const UmmaArgs& a, const TmaPair& tm, int nt, int planes, int src, int Z, cudaStream_t st) {
  // bf16x3 on tiles up to 128 wide: two MMAs per K step (A_hi x [B_hi | B_lo], A_lo x B_hi) instead of three -- 22 % / 17 %
  // fewer shared-memory operand bytes on 64- / 128-wide tiles; 256-wide tiles have no TMEM for a second accumulator block
  static const int concat = env_int("SAG_UMMA_CONCAT", 1);
  if (planes == 2) {
    if constexpr (BN <= 128) {
      if (concat) return launch_ns<BN, 2>(g, a, tm, nt, src, Z, st);
    }
    return launch_ns<BN, 3>(g, a, tm, nt, src, Z, st);
  }
  return launch_ns<BN, 1>(g, a, tm, nt, src, Z, st);
}

}  // namespace

// ---- host API -----------------------------------------------------------------------------------------------------
// 256-wide tiles halve the A re-reads of wide layers (the gather is L2-bandwidth bound); they use all 512 TMEM columns
// Tuning knob (development): SAG_UMMA_FORCE="MT:N:BN:Z,..." pins tile width / K split of the layers with MT m tiles
// and N columns (BN or Z = 0 keeps the planner's choice).
struct ForcedCfg { int64_t mt; int n, bn, z; };
static const std::vector<ForcedCfg>& forced_cfgs() {
  static std::vector<ForcedCfg> v;
  static bool init = false;
  if (!init) {
    init = true;
    const char* e = getenv("SAG_UMMA_FORCE");
    while (e != nullptr && *e) {
      long long mt = 0;
      int n = 0, bn = 0, z = 0;
      if (sscanf(e, "%lld:%d:%d:%d", &mt, &n, &bn, &z) == 4) v.push_back({(int64_t)mt, n, bn, z});
      e = strchr(e, ',');
      if (e) ++e;
    }
  }
  return v;
}
static const ForcedCfg* forced_for(int N, int64_t M) {
  for (const ForcedCfg& f : forced_cfgs())
    if (f.n == N && f.mt == cdiv64(M, UM_BM)) return &f;
  return nullptr;
}

// Tile width and K split of a contraction with K x N weights over M rows, from a cost model in microseconds fitted to
// per-layer CUDA-event timings on B200 (profiles/README.md, "planner"): a work item = its K chunks (shared-memory
// operand traffic grows with the tile width) + its epilogue passes; items run in waves over the 148 persistent CTAs;
// a K split adds the partial round trip through L2 and the reduce launch.  A single M tile (fully connected layers
// on a batch of windows) is weight-streaming bound: narrow tiles + a deep split spread the weights over all SMs.
struct TilePlan { int BN, Z; };
// Stream-K (see plan_streamk): on unless SAG_UMMA_STREAMK=0.  Measured on B200 at 32 windows: conv5_x 57.6 -> 47.5 us, conv4_x
// 53.6 -> 49.5 us, the forward 1880 -> 1937 audio-s/s.  (A first version lost: its raw-piece / fix-up epilogues ran serially at
// the end of every kernel without pipelined TMEM loads or prefetched slabs, and 256-wide pairs for conv5_x gained nothing.)
// Read at every call (tests switch it); a handle must be planned and run under the same setting.
static bool streamk_enabled() {
  const char* v = getenv("SAG_UMMA_STREAMK");
  return v == nullptr || atoi(v) != 0;
}
static TilePlan plan_tile(int K, int N, int64_t M) {
  static const int wide = env_int("SAG_UMMA_BN256", 1);
  static const int narrow_fc = env_int("SAG_UMMA_NARROW_FC", 1);
  static const int splitk = env_int("SAG_UMMA_SPLITK", 1);
  const int KC = cdiv(K, UM_BK);
  int cand[2], nc = 0;
  if (N <= 32) cand[nc++] = 32;
  else if (N <= 64 || (narrow_fc && M <= UM_BM)) cand[nc++] = 64;
  else { if (N >= 256 && wide) cand[nc++] = 256; cand[nc++] = 128; }      // the wide tile is the incumbent
  const ForcedCfg* f = forced_for(N, M);
  if (f && (f->bn == 32 || f->bn == 64 || f->bn == 128 || f->bn == 256)) { cand[0] = f->bn; nc = 1; }
  const int min_chunks = M <= UM_BM ? 2 : 4;      // a single M tile may split down to 2 chunks per item
  TilePlan best{cand[0], 1};
  double best_t = 0.0;
  for (int ci = 0; ci < nc; ++ci) {
    const int BN = cand[ci], NT = cdiv(N, BN);
    const int64_t tiles = cdiv64(M, UM_BM) * NT;
    const double chunk_us = BN == 256 ? 0.95 : (BN == 128 ? 0.59 : (BN == 64 ? 0.43 : 0.36));
    const double epi_us = 0.6 * (BN / 32);
    int z_lo = 1, z_hi = splitk ? std::min(32, KC / min_chunks) : 1;
    if (z_hi < 1) z_hi = 1;
    if (f && f->z >= 1 && f->z <= KC) z_lo = z_hi = f->z;
    int bz = z_lo;
    double bt = 0.0;
    for (int z = z_lo; z <= z_hi; ++z) {
      const int64_t waves = cdiv64(tiles * z, 148);
      double t = (double)waves * ((double)cdiv(KC, z) * chunk_us + epi_us) + 5.0;
      if (z > 1) t += 8.0 + (2.0 * z + 1.0) * (double)M * (double)(NT * BN) * 4.0 / 4.0e6;
      if (z == z_lo || t < bt * 0.97) { bt = t; bz = z; }     // split further only for a clear (>3 %) win
    }
    if (ci == 0 || bt < best_t * 0.95) { best_t = bt; best = TilePlan{BN, bz}; }
  }
  // Stream-K on CTA pairs with 256-wide tiles (plan_streamk below): the K chunks of the few wide tiles are dealt out over all
  // pairs, so the wide tile's halved weight traffic no longer costs whole idle waves (conv5_x at 32 windows: 26 pair tiles on 74
  // pairs).  Cost: the cluster's share of chunks at the pair rate (measured 1612 clk per 256 x 256 x 64 chunk) + the two
  // serial epilogues at the end (raw piece, then the fix-up) + launch.
  const int64_t MT = cdiv64(M, UM_BM);
  static const int sk_wide = env_int("SAG_UMMA_STREAMK_WIDE", 0);      // (measured: conv5_x 57.5 us on 256-wide pairs, 50 us on 128-wide tiles)
  if (sk_wide && streamk_enabled() && wide && !f && N >= 256 && KC >= 16 && MT >= 8) {
    const int64_t units = cdiv64(MT, 2) * cdiv(N, 256) * KC, G = max_conv_ctas() / 2;
    if (G >= 2 && units >= 4 * G) {
      const double t = (double)cdiv64(units, G) * 0.82 + 13.0 + 5.0;
      if (t < best_t * 0.9) best = TilePlan{256, 1};
    }
  }
  return best;
}

// Stream-K: when the tiles of an unsplit plan fill the last wave of the persistent grid badly (conv4_x at 32 windows: 49 tile
// pairs on 74 CTA pairs, conv5_x: 100 tiles on 148 CTAs -- a third of the SMs idle), the (tile, K chunk) units are dealt out
// evenly instead and the pieces of a tile meet in the epilogue of the CTA that finishes it (UmmaArgs::sk_flags).
static bool plan_streamk(int K, int N, int64_t M, const TilePlan& p) {
  const bool on = streamk_enabled();
  static const int thr = env_int("SAG_UMMA_STREAMK_EFF", 80);       // use it below this wave efficiency (percent)
  const int KC = cdiv(K, UM_BK);
  if (!on || p.Z != 1 || p.BN < 64 || KC < 8) return false;
  const int64_t MT = cdiv64(M, UM_BM);
  const int CL = (p.BN == 256 && KC >= 16 && MT >= 2) ? 2 : 1;       // (the pair rule of launch_ns)
  const int64_t T = cdiv64(MT, CL) * cdiv(N, p.BN), G = max_conv_ctas() / CL;
  if (G < 2 || T * KC < 4 * G) return false;
  const int64_t waves = cdiv64(T, G);
  return T * 100 < waves * G * thr;
}
static size_t streamk_slab_bytes(int BN) { return sizeof(float) * (size_t)num_sms() * UM_BM * (size_t)BN; }

// cuTensorMapEncodeTiled / cuTensorMapEncodeIm2col through the runtime's driver entry point lookup: libsag.so carries no
// link-time dependency on libcuda.so (it must load -- and report "no device" -- on machines without a driver)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// the packed weight image as a 2-D bf16 tensor of 128-byte rows, fetched in boxes of BN/2 rows (no swizzle: the image is
// already in the shared-memory layout)
static void make_weight_map(UmmaWeights* w) {
  static_assert(sizeof(CUtensorMap) == sizeof(w->wmap), "CUtensorMap size");
  w->wmap_ok = 0;
  const EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr || w->packed == nullptr || w->BN < 64) return;
  const cuuint64_t rows = (cuuint64_t)w->NT * w->KC * w->planes * w->BN;
  const cuuint64_t dims[2] = {64, rows};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)(w->BN / 2)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(w->wmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w->packed, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  w->wmap_ok = r == CUDA_SUCCESS ? 1 : 0;
}

void umma_free(UmmaWeights* w) {
  if (w->packed) cudaFree(w->packed);
  if (w->col_off) cudaFree(w->col_off);
  if (w->col_dy) cudaFree(w->col_dy);
  if (w->col_dx) cudaFree(w->col_dx);
  if (w->col_bias) cudaFree(w->col_bias);
  *w = UmmaWeights();
}

int umma_pack_weights(const float* wk, int K, int N, int64_t ldw, int precision, int64_t M, UmmaWeights* out, cudaStream_t st) {
  SAG_REQUIRE(precision == SAG_PREC_BF16 || precision == SAG_PREC_BF16X3, SAG_EUNSUPPORTED,
              "tcgen05 path: precision %d is not built (use bf16 or bf16x3)", precision);
  SAG_REQUIRE(K >= 0 && N > 0, SAG_EINVAL, "umma_pack_weights: bad shape %dx%d", K, N);
  UmmaWeights w;
  w.K = K; w.N = N;
  w.KC = cdiv(K, UM_BK);
  w.BN = plan_tile(K, N, M).BN;
  w.M_hint = M;
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
    make_weight_map(&w);
  }
  *out = w;
  return SAG_OK;
}

int umma_pack_conv_s2d(const float* w_hwio, int kh, int kw, int cin, int cout, int precision, int64_t M, int int_frames, UmmaWeights* out,
                       cudaStream_t st) {
  SAG_REQUIRE(4 * cin <= 16, SAG_EUNSUPPORTED, "space-to-depth route: %d input channels do not fit 16-channel pixels", cin);
  const int th = (kh + 1) / 2, tw = (kw + 1) / 2;
  float* wk = nullptr;
  const int64_t total = (int64_t)th * tw * 16 * cout;
  SAG_CHECK_CUDA(cudaMalloc(&wk, sizeof(float) * (size_t)total));
  int64_t blocks = cdiv64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  // int_frames: the frames arrive as 2k - 255 = 510 * (k/255 - 0.5): the weights carry the 1/510 (one correctly rounded division)
  s2d_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(w_hwio, kh, kw, cin, cout, th, tw, int_frames ? 510.f : 1.f, wk);
  int r = umma_pack_weights(wk, th * tw * 16, cout, cout, precision, M, out, st);
  if (r == SAG_OK && int_frames) out->a_single = 1;
  cudaStreamSynchronize(st);
  cudaFree(wk);
  return r;
}

// Sub-pixel formulation of tf.nn.conv2d_transpose VALID (core.py:139-140): every cell (u, v) of the
// (H+ty-1) x (W+tx-1) grid produces its sh x sw x Cout outputs from ty x tx taps of the input.
int umma_pack_deconv(const float* w_hwoi, const float* bias, int kh, int kw, int cout, int cin, int sh, int sw, int order,
                     int64_t y_sh, int64_t y_sw, int64_t y_sc, int precision, int64_t M, UmmaWeights* out, cudaStream_t st) {
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
  int r = umma_pack_weights(wk, K, N, N, precision, M, &w, st);
  cudaStreamSynchronize(st);
  cudaFree(wk);
  SAG_TRY(r);
  std::vector<int> off(N);
  std::vector<short> dy(N), dx(N);
  bool vec = true;
  for (int n = 0; n < N; ++n) {
    int py, px, co;
    decode_subpixel_column(order, n, cout, sw, &py, &px, &co);
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
  // groups of 8 columns that are 8 consecutive floats of one output row (planar outputs, stride 8 along the row): the
  // bulk-run store of the epilogue (out_mode 2)
  bool run8 = N % 32 == 0 && sw == 8 && y_sw == 1;
  for (int n = 0; n + 7 < N && run8; n += 8) {
    if (off[n] % 4 != 0) run8 = false;
    for (int e = 1; e < 8; ++e)
      if (off[n + e] != off[n] + e || dy[n + e] != dy[n] || dx[n + e] != dx[n] + e) run8 = false;
  }
  w.run8 = run8 ? 1 : 0;
  w.order = order;
  if (order == 2) { SAG_REQUIRE(cout == 32 && sw == 8 && w.BN == 256, SAG_EUNSUPPORTED, "deconv: column order 2 needs 32 channels, stride 8 and 256-wide tiles"); w.vec4 = 0; }
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
int umma_tile_width(int K, int N, int64_t M) { return plan_tile(K, N, M).BN; }

// Split-K plan: layers whose tile count cannot fill the 148 SMs but whose K loop is long are cut along K; the partial
// accumulators go through `scratch` ([Z][M][n_pad] fp32) and splitk_reduce_kernel finishes them (deterministic: fixed
// summation order).  (A fix-up by the last CTA to arrive at a tile was measured 25 % slower end to end: 128 threads
// pulling Z partial tiles through L2 latency cannot compete with a reduce spread over the whole GPU.)
int umma_split_k(int K, int N, int64_t M, size_t* scratch_bytes) {
  const TilePlan p = plan_tile(K, N, M);
  // partial slabs are padded to whole 128-row tiles (the TMA store of a tile never crosses into the next slab)
  if (scratch_bytes) *scratch_bytes = p.Z > 1 ? sizeof(float) * (size_t)p.Z * (size_t)(cdiv64(M, UM_BM) * UM_BM) * (size_t)(cdiv(N, p.BN) * p.BN) : 0;
  if (scratch_bytes && plan_streamk(K, N, M, p)) *scratch_bytes = streamk_slab_bytes(p.BN);      // (stream-K plans have Z == 1)
  return p.Z;
}

bool umma_stream_k(int K, int N, int64_t M) { return plan_streamk(K, N, M, plan_tile(K, N, M)); }

int streamk_schedule(int64_t tiles, int kc, int clusters, int cluster, int* items, int max_items) {
  SkRange r;
  r.init(tiles, kc, clusters, cluster);
  for (int j = 0; j < r.n && j < max_items; ++j) {
    int64_t t;
    int kb, ke, ff;
    bool produce;
    r.item(j, &t, &kb, &ke, &produce, &ff);
    items[5 * j] = (int)t; items[5 * j + 1] = kb; items[5 * j + 2] = ke; items[5 * j + 3] = produce ? 1 : 0; items[5 * j + 4] = ff;
  }
  return r.n;
}

thread_local int g_umma_tma = -1;   // -1: SAG_UMMA_TMA (default on); 0 / 1: forced (sag_set_option "tma_gather")
thread_local int g_umma_pair = -1;  // -1: SAG_UMMA_PAIR (default off); 0 / 1: forced (sag_set_option "cta_pair")

// im2col tensor maps over the two bf16 planes of an NHWC activation for the geometry's taps (reference for the
// corner arithmetic: base pixel of output (i, j) = (i*isy + min dy, j*isx + min dx); the box's upper corner is chosen so
// that exactly PW x PH base pixels fit).  Returns false when the layer does not fit the TMA unit's limits.
// cuTensorMapEncodeIm2col through the runtime's driver entry point lookup: libsag.so carries no link-time dependency
// on libcuda.so (it must load -- and report "no device" -- on machines without a driver)
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeIm2colFn>(p);
  }();
  return fn;
}

static bool make_im2col_maps(const ActView& x, const GatherGeom& g, TmaPair* tm, int* w0, int* h0) {
  if (x.fmt != ACT_BF2 || g.Cin % UM_BK != 0 || g.x_ld % 8 != 0 || g.T < 1) return false;
  if ((reinterpret_cast<uintptr_t>(x.p) & 15) != 0 || (x.plane & 15) != 0) return false;
  int mindx = g.dx[0], mindy = g.dy[0], maxdx = g.dx[0], maxdy = g.dy[0];
  for (int t = 1; t < g.T; ++t) {
    mindx = std::min(mindx, (int)g.dx[t]); maxdx = std::max(maxdx, (int)g.dx[t]);
    mindy = std::min(mindy, (int)g.dy[t]); maxdy = std::max(maxdy, (int)g.dy[t]);
  }
  const int up_w = (g.PW - 1) * g.isx + 1 + mindx - g.W, up_h = (g.PH - 1) * g.isy + 1 + mindy - g.H;
  if (mindx < -128 || mindx > 127 || mindy < -128 || mindy > 127 || up_w < -128 || up_w > 127 || up_h < -128 || up_h > 127) return false;
  if (maxdx - mindx > 255 || maxdy - mindy > 255 || g.isx < 1 || g.isx > 8 || g.isy < 1 || g.isy > 8) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
  const cuuint64_t x_row = g.x_row != 0 ? (cuuint64_t)g.x_row : (cuuint64_t)g.W * g.x_ld;
  const cuuint64_t strides[3] = {(cuuint64_t)g.x_ld * 2, x_row * 2, (cuuint64_t)g.H * x_row * 2};
  const int lower[2] = {mindx, mindy}, upper[2] = {up_w, up_h};
  const cuuint32_t estr[4] = {1, (cuuint32_t)g.isx, (cuuint32_t)g.isy, 1};
  static int driver = -1;
  if (driver < 0) cudaDriverGetVersion(&driver);
  const EncodeIm2colFn encode = encode_im2col_fn();
  if (encode == nullptr) return false;
  for (int pl = 0; pl < (x.plane != 0 ? 2 : 1); ++pl) {
    CUtensorMap* m = pl == 0 ? &tm->hi : &tm->lo;
    void* base = reinterpret_cast<char*>(x.p) + (pl == 0 ? 0 : x.plane);
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, lower, upper,
                                         (cuuint32_t)UM_BK, (cuuint32_t)UM_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    // same adjustment CUTLASS applies to im2col maps of tensors under 128 KB on drivers up to 13.1
    // (cute/atom/copy_traits_sm90_im2col.hpp, make_im2col_tma_copy_desc)
    if (driver <= 13010 && (cuuint64_t)g.N * strides[2] < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  *w0 = mindx;
  *h0 = mindy;
  return true;
}

// ---- halo-resident 3x3 convolution: eligibility, tile shape, tensor maps, launch -----------------------------------------
thread_local int g_umma_halo = -1;  // -1: SAG_UMMA_HALO (default on); 0 / 1: forced (sag_set_option "halo_conv")

template <int BN, bool PAIR>
static int launch_halo(const HaloArgs& a, const HaloMaps& tm, size_t a_slot, cudaStream_t st) {
  auto kern = halo_conv_umma_kernel<BN, PAIR>;
  static int budget[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (budget[dev & 63] == 0) {
    cudaFuncAttributes fa;
    int optin = 0;
    SAG_CHECK_CUDA(cudaFuncGetAttributes(&fa, kern));
    SAG_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int b = optin - (int)fa.sharedSizeBytes - env_int("SAG_UMMA_SMEM_RESERVE", 0);   // (development: shared memory left to co-resident CTAs of other kernels)
    SAG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    budget[dev & 63] = b;
  }
  constexpr int B_BYTES = (BN + (PAIR ? BN / 2 : BN)) * 128;
  constexpr int CL = PAIR ? 2 : 1;
  const size_t fixed = 256 + 1024 + 32768;
  HaloArgs args = a;
  const int taps = a.NDX * a.NDY * a.CC;
  // weights resident when every tap fits beside two activation boxes (one N tile: the same chunks serve every tile of the CTA)
  static const int bres_env = env_int("SAG_UMMA_HALO_BRES", 1);
  static const int halo_dbg = env_int("SAG_HALO_DEBUG", 0);
  args.dbg = halo_dbg;
  static const int fast_taps = env_int("SAG_HALO_FAST_TAPS", 1);
  args.fast_taps = fast_taps;
  args.BRES = (bres_env && a.NT == 1 && taps <= HL_MAX_SB &&
               (long)fixed + 2 * (long)a_slot + (long)taps * B_BYTES <= (long)budget[dev & 63]) ? 1 : 0;
  if (args.BRES) {
    args.SB = taps;
    long sa = ((long)budget[dev & 63] - (long)fixed - (long)taps * B_BYTES) / (long)a_slot;
    args.SA = sa > HL_MAX_SA ? HL_MAX_SA : (int)sa;
  } else {
    args.SA = 3;
    long sb = ((long)budget[dev & 63] - (long)fixed - 3 * (long)a_slot) / B_BYTES;
    if (sb < 2) { args.SA = 2; sb = ((long)budget[dev & 63] - (long)fixed - 2 * (long)a_slot) / B_BYTES; }
    SAG_REQUIRE(sb >= 2, SAG_EUNSUPPORTED, "halo conv: the tile does not fit the shared memory");
    args.SB = sb > HL_MAX_SB ? HL_MAX_SB : (int)sb;
  }
  const size_t smem = fixed + (size_t)args.SA * a_slot + (size_t)args.SB * B_BYTES;
  const int tiles = a.NIMG * a.TY * a.TX;
  const int n_work = cdiv(tiles, CL) * a.NT;
  int clusters = max_conv_ctas() / CL;
  if (clusters < 1) clusters = 1;
  if (n_work < clusters) clusters = n_work;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  cfg.blockDim = dim3(HL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  memset(attrs, 0, sizeof(attrs));
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CL;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // debug: SAG_HALO_TRACE=<taps> prints where the three roles of the first launches with that many taps wait
  static const int trace_taps = env_int("SAG_HALO_TRACE", 0);
  static int traced = 0;
  if (trace_taps > 0 && taps == trace_taps && traced < 2) {
    ++traced;
    const size_t n = (size_t)cfg.gridDim.x * 8;
    long long* dtr = nullptr;
    cudaMalloc(&dtr, n * sizeof(long long));
    cudaMemset(dtr, 0, n * sizeof(long long));
    args.trace = dtr;
    cudaStreamSynchronize(st);
    cudaLaunchKernelEx(&cfg, kern, args, tm);
    cudaStreamSynchronize(st);
    std::vector<long long> tr(n);
    cudaMemcpy(tr.data(), dtr, n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dtr);
    double d[8] = {0};
    int lead = 0;
    for (size_t c = 0; c < n / 8; ++c) {
      for (int i = 0; i < 8; ++i) d[i] += (double)tr[c * 8 + i];
      if (tr[c * 8 + 2] > 0) ++lead;
    }
    const double nc = (double)(n / 8), nl = lead > 0 ? (double)lead : 1.0;
    fprintf(stderr, "[halo trace] BN=%d pair=%d ctas=%u tiles=%d taps=%d SA=%d SB=%d BRES=%d A1=%d\n", BN, (int)PAIR, cfg.gridDim.x, tiles, taps, args.SA,
            args.SB, args.BRES, args.A1);
    fprintf(stderr, "[halo trace]   producer : total %9.0f clk, waiting for a free box %9.0f\n", d[0] / nc, d[1] / nc);
    fprintf(stderr, "[halo trace]   MMA      : total %9.0f clk, waiting for a box %9.0f, for a free accumulator %9.0f\n", d[2] / nl, d[3] / nl, d[4] / nl);
    fprintf(stderr, "[halo trace]   epilogue : total %9.0f clk, waiting for an accumulator %9.0f\n", d[5] / nc, d[6] / nc);
    SAG_LAUNCH_CHECK();
    return SAG_OK;
  }
  SAG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args, tm));
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

// returns SAG_OK and sets *done when the layer ran on the halo kernel; *done == false: not eligible, take the im2col kernel
static int try_halo_conv(const ActView& x, const UmmaWeights& w, const ActView& y, const GatherGeom& g, const Epilogue& ep, int Z,
                         cudaStream_t st, bool* done) {
  *done = false;
  // -1 (default): on where no tile overhangs the image; 1: forced wherever the layer is eligible (tests); 0: off
  static const int halo_env = env_int("SAG_UMMA_HALO", -1);
  const int halo_want = g_umma_halo >= 0 ? g_umma_halo : halo_env;
  if (halo_want == 0) return SAG_OK;
  if (Z != 1 || g.isy != 1 || g.isx != 1 || g.Cin % 64 != 0 || g.x_ld % 8 != 0) return SAG_OK;
  // taps must form a full NDX x NDY rectangle in row-major order (3 x 3 SAME convolutions: dx, dy in -1..1; ResNet conv1 in its
  // space-to-depth form: one column of four rows, the four horizontal taps being the overlapping 64-channel pixel)
  int ndx = 0, ndy = 0;
  const int dx0 = g.dx[0], dy0 = g.dy[0];
  for (int t = 0; t < g.T && g.dy[t] == dy0; ++t) ++ndx;
  if (ndx < 1 || g.T % ndx != 0) return SAG_OK;
  ndy = g.T / ndx;
  for (int t = 0; t < g.T; ++t)
    if (g.dy[t] != dy0 + t / ndx || g.dx[t] != dx0 + t % ndx) return SAG_OK;
  if (ndy < 2 || ndx > 3 || ndy > 4) return SAG_OK;
  // conv1 (one tap column, four rows): its time is set by the 205 MB of raw fp32 output, not by operand bytes -- measured 125.6 us
  // on the im2col kernel, 127-136 us here -- so it takes this kernel only when forced (tests)
  static const int conv1_env = env_int("SAG_UMMA_HALO_CONV1", 1);
  if (ndx == 1 && halo_want <= 0 && !(halo_want < 0 && conv1_env)) return SAG_OK;
  const bool a1 = w.a_single != 0;                   // one exact activation plane (integer frames)
  if (x.fmt != ACT_BF2 || (x.plane == 0) != a1 || (reinterpret_cast<uintptr_t>(x.p) & 15) != 0 || (x.plane & 15) != 0) return SAG_OK;
  if (w.planes != 2 || (w.BN != 64 && w.BN != 128) || w.N > 512 || w.col_off != nullptr || w.K != g.T * g.Cin) return SAG_OK;
  if (ep.bias != nullptr || ep.relu || y.fmt != ACT_F32 || g.y_sc != 1 || g.y_sw % 4 != 0 || g.oy0 != 0 || g.ox0 != 0 || g.osy != 1 || g.osx != 1 ||
      g.y_sh != (int64_t)g.PW * g.y_sw || g.y_sn != (int64_t)g.PH * g.y_sh || (reinterpret_cast<uintptr_t>(y.p) & 15) != 0)
    return SAG_OK;
  const int OH = g.PH, OW = g.PW;                    // output grid (== the image for SAME 3 x 3; the padded image is larger for conv1)
  // tile shape: least padding, then the tallest (fewest halo rows per output row).  The activation bytes saved must outweigh
  // the padded tiles: layers with small images (conv4_x: 14 x 28, conv5_x: 7 x 14) stay on the im2col kernel, whose tiles
  // run across image boundaries.
  int TW = 0, TH = 0;
  double best = 1e30;
  for (int tw = 8; tw <= 64; tw *= 2) {
    const int th = UM_BM / tw;
    const double padded = (double)(cdiv(OW, tw) * tw) * (cdiv(OH, th) * th) / ((double)OW * OH);
    const double score = padded * (1.0 + 0.01 * (th + ndy - 1.0) / th);
    if (score < best) { best = score; TW = tw; TH = th; }
  }
  const double padded = (double)(cdiv(OW, TW) * TW) * (cdiv(OH, TH) * TH) / ((double)OW * OH);
  if (padded > (halo_want > 0 ? 1.3 : 1.001)) return SAG_OK;      // measured: conv3_x with 14 % padded tiles loses 20 % (profiles/README.md); conv2_x (none) gains
  const EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr) return SAG_OK;
  HaloMaps tm;
  memset(&tm, 0, sizeof(tm));
  {
    // (x_row != 0: pixels overlap -- conv1's 64-channel pixels at a 16-channel pitch, see resnet18_tower)
    const cuuint64_t x_row = g.x_row != 0 ? (cuuint64_t)g.x_row : (cuuint64_t)g.W * g.x_ld;
    const cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    const cuuint64_t strides[3] = {(cuuint64_t)g.x_ld * 2, x_row * 2, (cuuint64_t)g.H * x_row * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)(TH + ndy - 1), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int pl = 0; pl < (a1 ? 1 : 2); ++pl) {
      void* base = reinterpret_cast<char*>(x.p) + (pl == 0 ? 0 : x.plane);
      if (encode(pl == 0 ? &tm.hi : &tm.lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return SAG_OK;
    }
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)w.N, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)g.N};
    const cuuint64_t strides[3] = {(cuuint64_t)g.y_sw * 4, (cuuint64_t)g.y_sh * 4, (cuuint64_t)g.y_sn * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode(&tm.o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return SAG_OK;
  }
  HaloArgs a;
  memset(&a, 0, sizeof(a));
  a.wpacked = reinterpret_cast<const uint8_t*>(w.packed);
  a.stat_sum = ep.stat_sum; a.stat_sqs = ep.stat_sqs;
  a.NIMG = g.N; a.H = OH; a.W = OW; a.TW = TW; a.TH = TH; a.TX = cdiv(OW, TW); a.TY = cdiv(OH, TH);
  a.CC = g.Cin / 64; a.Ntot = w.N; a.NT = w.NT;
  a.NDX = ndx; a.NDY = ndy; a.DX0 = dx0; a.DY0 = dy0;
  a.A1 = a1 ? 1 : 0;
  const size_t a_slot = (a1 ? 1 : 2) * (size_t)(TH + ndy - 1) * TW * 128;
  if (34048 + 2 * a_slot + 2 * (size_t)(2 * w.BN * 128) > 220 * 1024) return SAG_OK;      // two boxes + two weight chunks must fit
  // CTA pairs (two neighbouring tiles per cluster, half the weight bytes and half the MMA instructions per tile): default on
  static const int pair_env = env_int("SAG_UMMA_PAIR", -1);
  const int want = g_umma_pair >= 0 ? g_umma_pair : pair_env;
  static const int conv1_pair_env = env_int("SAG_UMMA_HALO_CONV1_PAIR", 0);     // measured: 90.3 us alone, 107.2 us on pairs (resident weights: nothing left to share)
  const bool pair = (want < 0 || want != 0) && w.wmap_ok && g.N * a.TY * a.TX >= 2 && !(ndx == 1 && want < 0 && !conv1_pair_env);
  if (pair) {
    memcpy(&tm.w, w.wmap, sizeof(tm.w));
    SAG_TRY(w.BN == 64 ? (launch_halo<64, true>(a, tm, a_slot, st)) : (launch_halo<128, true>(a, tm, a_slot, st)));
  } else {
    SAG_TRY(w.BN == 64 ? (launch_halo<64, false>(a, tm, a_slot, st)) : (launch_halo<128, false>(a, tm, a_slot, st)));
  }
  *done = true;
  return SAG_OK;
}

bool umma_int_frames_supported(int n, int oh, int ow) {
  static const int on = env_int("SAG_UMMA_INT_FRAMES", 1);
  static const int halo_env = env_int("SAG_UMMA_HALO", -1);
  static const int conv1_env = env_int("SAG_UMMA_HALO_CONV1", 1);
  const int halo_want = g_umma_halo >= 0 ? g_umma_halo : halo_env;
  if (!on || halo_want == 0 || (halo_want < 0 && !conv1_env) || encode_tiled_fn() == nullptr || n < 1) return false;
  for (int tw = 8; tw <= 64; tw *= 2)                 // a tile shape without padded tiles (try_halo_conv's default rule)
    if (ow % tw == 0 && oh % (UM_BM / tw) == 0 && 34048 + 2 * (size_t)(UM_BM / tw + 3) * tw * 128 + 4 * 16384 <= 220 * 1024) return true;
  return false;
}

int launch_gather_gemm_umma(const ActView& x, const UmmaWeights& w, const ActView& y, const GatherGeom& g, const Epilogue& ep,
                            int oh_lim, int ow_lim, float* scratch, cudaStream_t st) {
  SAG_REQUIRE(w.packed != nullptr || w.KC == 0, SAG_ESTATE, "tcgen05 path: weights are not packed");
  SAG_REQUIRE(g.T * g.Cin == w.K, SAG_EINVAL, "tcgen05 path: geometry K %d does not match packed K %d", g.T * g.Cin, w.K);
  const int64_t M = (int64_t)g.N * g.PH * g.PW;
  if (M == 0) return SAG_OK;
  SAG_REQUIRE(M < (1ll << 31), SAG_EINVAL, "tcgen05 path: too many rows");
  UmmaArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x.p; a.x_plane = x.plane;
  a.wpacked = reinterpret_cast<const uint8_t*>(w.packed);
  a.y = y.p; a.y_plane = y.plane;
  a.out_bf2 = y.fmt == ACT_BF2 ? (y.plane != 0 ? 2 : 1) : 0;
  a.relu = ep.relu; a.stat_sum = ep.stat_sum; a.stat_sqs = ep.stat_sqs;
  a.col_off = w.col_off; a.col_dy = w.col_dy; a.col_dx = w.col_dx;
  a.bias = w.col_off != nullptr ? w.col_bias : ep.bias;
  a.oh_lim = oh_lim; a.ow_lim = ow_lim;
  a.Ntot = w.N; a.K = w.K; a.KC = w.KC;
  a.n_pad = w.NT * w.BN;
  const bool aligned_y = (reinterpret_cast<uintptr_t>(y.p) & 15) == 0 && (y.plane & 15) == 0;
  if (w.col_off != nullptr) {
    SAG_REQUIRE(ep.stat_sum == nullptr, SAG_EUNSUPPORTED, "tcgen05 path: statistics with mapped outputs");
    a.vec_store = (w.vec4 && aligned_y && g.y_sn % 4 == 0 && g.y_sh % 4 == 0 && (g.y_sw * g.osx) % 4 == 0 && ow_lim % 4 == 0) ? 1 : 0;
  } else {
    a.vec_store = (g.y_sc == 1 && aligned_y && w.N % 4 == 0 && g.y_sn % 4 == 0 && g.y_sh % 4 == 0 && g.y_sw % 4 == 0) ? 1 : 0;
  }
  // vector gather: every 8-element K group is one tap's 8 contiguous, 16-byte aligned elements
  const int elt = x.fmt == ACT_BF2 ? 8 : 4;     // elements per 16 bytes
  bool vec = (g.Cin % 8 == 0) && ((g.x_ld * g.isx) % elt == 0) && ((g.x_row != 0 ? g.x_row : g.x_ld * g.W) % elt == 0) &&
             ((reinterpret_cast<uintptr_t>(x.p) & 15) == 0) && (x.plane & 15) == 0;
  for (int t = 0; t < g.T && vec; ++t) vec = (g.dx[t] * g.x_ld) % elt == 0;
  int src = vec ? SRC_F32_VEC : SRC_F32;
  if (x.fmt == ACT_BF2) {
    SAG_REQUIRE(vec, SAG_EUNSUPPORTED, "tcgen05 path: split-bf16 activations need 16-byte aligned 8-channel groups (Cin %d, ld %lld)",
                g.Cin, (long long)g.x_ld);
    SAG_REQUIRE(w.planes == 1 || x.plane != 0 || w.a_single, SAG_EINVAL, "tcgen05 path: bf16x3 needs the lo plane of the activation");
    src = SRC_BF2;
  }
  TmaPair tm;
  memset(&tm, 0, sizeof(tm));
  memcpy(&tm.w, w.wmap, sizeof(tm.w));
  a.pair_ok = w.wmap_ok;
  static const int tma_env = env_int("SAG_UMMA_TMA", 1);
  if (src == SRC_BF2 && w.BN >= 64 && (g_umma_tma < 0 ? tma_env : g_umma_tma) && make_im2col_maps(x, g, &tm, &a.tma_w0, &a.tma_h0))
    src = SRC_TMA;
  const TilePlan plan = plan_tile(w.K, w.N, M);
  SAG_REQUIRE(w.BN == plan.BN, SAG_ESTATE, "tcgen05 path: weights were packed for a different row count (tile %d, planned %d)", w.BN, plan.BN);
  const int Z = scratch != nullptr ? plan.Z : 1;
  {
    bool done = false;
    SAG_TRY(try_halo_conv(x, w, y, g, ep, Z, st, &done));
    if (done) return SAG_OK;
  }
  SAG_REQUIRE(!w.a_single, SAG_EUNSUPPORTED, "tcgen05 path: single-plane integer frames run on the halo kernel only (umma_int_frames_supported)");
  static const int epi_dbg = env_int("SAG_UMMA_EPI_DEBUG", 0);
  a.dbg = epi_dbg;
  if (Z > 1) a.partial = scratch;
  static const int sk_mapped = env_int("SAG_UMMA_STREAMK_DECONV", 0);      // sub-pixel transposed convs (mapped outputs): measured deconv3 29.1 -> 33.3 us, others +-0 -- their epilogues pace them; off
  if (Z == 1 && scratch != nullptr && ep.sk_flags != nullptr && (w.col_off == nullptr || sk_mapped) && ep.gains == nullptr &&
      plan_streamk(w.K, w.N, M, plan)) {
    a.sk_flags = ep.sk_flags;
    a.sk_slab = scratch;
  }
  // output pixel m sits at element m*y_sw: the epilogue needs no (n, i, j) decode
  a.dense = (w.col_off == nullptr && g.osy == 1 && g.osx == 1 && g.oy0 == 0 && g.ox0 == 0 &&
             g.y_sh == (int64_t)g.PW * g.y_sw && g.y_sn == (int64_t)g.PH * g.y_sh) ? 1 : 0;
  a.m_pad = cdiv64(M, UM_BM) * UM_BM;
  // how the epilogue's passes leave the CTA (see UmmaArgs::out_mode)
  static const int tma_out_env = env_int("SAG_UMMA_TMA_STORE", 1);
  a.out_mode = 0;
  if (tma_out_env) {
    const EncodeTiledFn encode = encode_tiled_fn();
    const bool raw = Z > 1;
    if (encode != nullptr && (raw || (a.dense && a.vec_store && a.out_bf2 == 0 && w.col_off == nullptr))) {
      const cuuint64_t dims[2] = {(cuuint64_t)(raw ? a.n_pad : w.N), (cuuint64_t)(raw ? (int64_t)Z * a.m_pad : M)};
      const cuuint64_t strides[1] = {(cuuint64_t)(raw ? a.n_pad : g.y_sw) * 4};
      const cuuint32_t box[2] = {32, (cuuint32_t)UM_BM};
      const cuuint32_t estr[2] = {1, 1};
      void* base = raw ? (void*)scratch : y.p;
      if ((reinterpret_cast<uintptr_t>(base) & 15) == 0 && strides[0] % 16 == 0 &&
          encode(&tm.o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        a.out_mode = 1;
    } else if (ep.gains != nullptr) {
      // (checked below: the fusion has no fallback inside this launch -- the caller picks the unfused path instead)
    } else if (!raw && w.col_off != nullptr && w.run8 && a.out_bf2 == 0 && g.PW == UM_BM && g.ox0 == 0 && g.osx == 8 && g.y_sw == 1 &&
               ow_lim == UM_BM * 8 && aligned_y && g.y_sn % 4 == 0 && g.y_sh % 4 == 0) {
      a.out_mode = 2;
    }
  }
  if (ep.gains != nullptr) {
    SAG_REQUIRE(w.order == 2 && w.BN == 256 && src == SRC_TMA && Z == 1 && g.PW == UM_BM && g.ox0 == 0 && g.osx == 8 &&
                ow_lim == UM_BM * 8 && w.N % 256 == 0 && ep.gain_loc != nullptr && w.col_bias != nullptr &&
                (reinterpret_cast<uintptr_t>(ep.gains) & 15) == 0 && ep.gain_plane % 4 == 0,
                SAG_EUNSUPPORTED, "tcgen05 path: this launch cannot fuse the mask gains");
    a.out_mode = 3;
    a.gain_loc = ep.gain_loc; a.gains = ep.gains; a.gain_plane = ep.gain_plane;
  }
  SAG_REQUIRE(w.KC >= 1, SAG_EINVAL, "tcgen05 path: empty contraction");
  SAG_REQUIRE(ep.stat_sum == nullptr || w.N <= UM_MAX_N, SAG_EUNSUPPORTED, "tcgen05 path: statistics over %d columns", w.N);
  SAG_REQUIRE(ep.stat_sum == nullptr || Z == 1 || 256 % cdiv(w.N, 4) == 0, SAG_EUNSUPPORTED, "tcgen05 path: split-K statistics over %d columns", w.N);
  int r;
  switch (w.BN) {
    case 32: r = launch_bn<32>(g, a, tm, w.NT, w.planes, src, Z, st); break;
    case 64: r = launch_bn<64>(g, a, tm, w.NT, w.planes, src, Z, st); break;
    case 128: r = launch_bn<128>(g, a, tm, w.NT, w.planes, src, Z, st); break;
    case 256: r = launch_bn<256>(g, a, tm, w.NT, w.planes, src, Z, st); break;
    default: set_error("tcgen05 path: unsupported tile width %d", w.BN); return SAG_EINVAL;
  }
  SAG_TRY(r);
  if (Z > 1) {
    const int rpb = reduce_rows_per_block(M);
    launch_pdl(splitk_reduce_kernel, dim3((unsigned)cdiv64(M, rpb)), dim3(256), 0, st, g, a, Z, rpb);
    SAG_LAUNCH_CHECK();
  }
  return SAG_OK;
}

}  // namespace sag
