// Micro-benchmark: issue rate of tcgen05.mma kind::f16, M = 128, K = 16, by N, operand form (A in TMEM or in shared
// memory) and accumulator pattern (same tile every time, or rotating over several tiles).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// form 0 = TS, 1 = SS; nrot = number of accumulator tiles rotated over; kchain = consecutive MMAs on one tile
__global__ void k_bench(int N, int form, int nrot, int kchain, int iters, long long* cyc) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm + 16384);
    const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(((a_addr >> 4) & 0x3FFF) | (1u << 16));
    const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_addr >> 4) & 0x3FFF) | (1u << 16));
    const uint32_t a_tmem = tb + 448;     // 32 packed columns of A at the top of TMEM
    const long long t0 = clock64();
    int tile = 0;
    for (int it = 0; it < iters; ++it) {
      for (int ks = 0; ks < kchain; ++ks) {
        const uint32_t d = tb + (uint32_t)(tile * N);
        if (form == 0) mma_ts(d, a_tmem + (uint32_t)(ks & 3) * 8u, db + (uint64_t)((ks & 3) * 2), idesc, ks ? 1u : 0u);
        else mma_ss(d, da + (uint64_t)((ks & 3) * 2), db + (uint64_t)((ks & 3) * 2), idesc, ks ? 1u : 0u);
      }
      tile = (tile + 1) % nrot;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cyc = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 8);
  const size_t smem = 64 * 1024 + 1024;
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Cfg { int N, form, nrot, kchain; };
  const Cfg cfgs[] = {{64, 0, 1, 4}, {64, 0, 2, 4}, {64, 0, 6, 4}, {64, 0, 6, 1}, {128, 0, 1, 4}, {128, 0, 3, 4}, {128, 0, 3, 1},
                      {256, 0, 1, 4}, {64, 1, 1, 4}, {64, 1, 6, 4}, {128, 1, 1, 4}, {128, 1, 3, 4}, {256, 1, 1, 4}, {256, 1, 1, 1}};
  for (const Cfg& c : cfgs) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) k_bench<<<148, 128, smem>>>(c.N, c.form, c.nrot, c.kchain, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)iters * c.kchain);
    printf("N=%3d %s rot %d chain %d: %.1f cycles/MMA  (ideal %d)  -> %.0f FMA/clk/SM  %s\n", c.N, c.form ? "SS" : "TS", c.nrot, c.kchain,
           per, c.N / 2, 128.0 * c.N * 16 / per, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
