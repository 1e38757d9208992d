// tcgen05.mma issue microbenchmark (development tool, not part of libsag.so): how long does a chain of kind::f16 MMAs of shape
// M=128 x N x K=16 take per instruction on one SM, as a function of N (64 / 128 / 256) and of whether consecutive MMAs
// accumulate into the SAME TMEM accumulator (dependent) or alternate between 2 / 4 accumulators?  Operands: zero-filled
// SWIZZLE_128B K-major tiles in shared memory (the numbers do not matter), one elected thread issues, one commit at the end.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_microbench tools/umma_microbench.cu && ./umma_microbench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// N: MMA width; SETS: accumulators the chain alternates between.  The issue loop is unrolled 16x with compile-time operands so
// that it costs a few instructions per MMA (a first version with runtime modulo arithmetic measured its own loop: 144 clk).
template <int N, int SETS>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  // A: 4 chunks of 128 rows x 128 B; B: 4 chunks of 256 rows x 128 B
  const uint32_t a_base = base, b_base = base + 4 * 16384;
  for (uint32_t i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint64_t DESC_HI = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);   // LBO, SBO = 1024 B, version, SWIZZLE_128B
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint64_t da[4], db[4];
    for (int c = 0; c < 4; ++c) {
      da[c] = DESC_HI | (uint64_t)(((a_base + c * 16384) & 0x3FFFFu) >> 4);
      db[c] = DESC_HI | (uint64_t)(((b_base + c * 32768) & 0x3FFFFu) >> 4);
    }
    // first touch of every accumulator (accumulate = 0), outside the timed loop
    for (int s = 0; s < SETS; ++s) umma_bf16(tmem_base + s * N, da[0], db[0], idesc, 0u);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j)      // chunk j / 4, K step j % 4 of the chunk: like the convolution kernels
        umma_bf16(tmem_base + (uint32_t)((j % SETS) * N), da[j >> 2] + 2 * (j & 3), db[j >> 2] + 2 * (j & 3), idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int N, int SETS>
static void run(int grid, int iters, long long* d) {
  const size_t smem = 4 * 16384 + 4 * 32768 + 1024;
  auto k = bench_kernel<N, SETS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<grid, 128, smem>>>(iters, d);      // warm-up
  k<<<grid, 128, smem>>>(iters, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
  long long h[2 * 148];
  cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
  double issue = 0, done = 0;
  for (int c = 0; c < grid; ++c) { issue += (double)h[2 * c] / grid; done += (double)h[2 * c + 1] / grid; }
  printf("  ctas %3d  N=%3d  accumulators=%d : issue %6.1f  complete %6.1f   (tensor time %d clk)\n", grid, N, SETS, issue / iters,
         done / iters, N / 2);
}

int main() {
  long long* d = nullptr;
  cudaMalloc(&d, sizeof(long long) * 2 * 148);
  const int iters = 4096;
  printf("tcgen05.mma kind::f16 M=128 K=16, %d instructions by one thread; clk per instruction (issue loop / until the commit lands)\n", iters);
  for (int grid : {1, 148}) {
    run<64, 1>(grid, iters, d); run<64, 2>(grid, iters, d); run<64, 4>(grid, iters, d);
    run<128, 1>(grid, iters, d); run<128, 2>(grid, iters, d); run<128, 4>(grid, iters, d);
    run<256, 1>(grid, iters, d); run<256, 2>(grid, iters, d);
  }
  cudaFree(d);
  return 0;
}
