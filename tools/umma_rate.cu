// Issue rate of tcgen05.mma (kind::f16, bf16, M = 128, cta_group::1) for different shared-memory operand layouts: how many clocks
// one MMA takes when REPS of them are issued back to back by one thread (operands are whatever is in shared memory; only time matters).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate tools/umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}

__global__ void rate_kernel(int N, int a_major, int b_major, uint32_t layout, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                            uint32_t a_step, uint32_t b_step, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(smem), b0 = a0 + 96 * 1024, barp = (uint32_t)__cvta_generic_to_shared(&bar);
    const long long t0 = clock64();
    const uint64_t ad0 = desc(a0, a_lbo, a_sbo, layout), bd0 = desc(b0, b_lbo, b_sbo, layout);
    const uint64_t as16 = a_step >> 4, bs16 = b_step >> 4;
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {                 // descriptors differ in the start-address field only: one 64-bit add each
        const uint64_t ad = ad0 + (uint64_t)u * as16, bd = bd0 + (uint64_t)u * bs16;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad),
                     "l"(bd), "r"(idesc), "r"((r | u) > 0 ? 1u : 0u)
                     : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(barp) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24) && !ok; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(barp), "r"(0) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    out[2] = ok;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  struct Cfg { const char* name; int N, am, bm; uint32_t layout, albo, asbo, blbo, bsbo, astep, bstep; };
  const Cfg cfgs[] = {
      {"K-major no-swizzle, conv tile (LBO 2880, SBO 160), N=48", 48, 0, 0, 0, 2880, 160, 768, 128, 16, 4608},
      {"K-major no-swizzle, dense (LBO 2048, SBO 128), N=48", 48, 0, 0, 0, 2048, 128, 768, 128, 16, 4608},
      {"K-major no-swizzle, dense core matrices adjacent (LBO 128, SBO 256), N=48", 48, 0, 0, 0, 128, 256, 128, 256, 4096, 4608},
      {"K-major no-swizzle, conv tile, N=16", 16, 0, 0, 0, 2880, 160, 256, 128, 16, 4608},
      {"K-major no-swizzle, conv tile, N=160", 160, 0, 0, 0, 2880, 160, 2560, 128, 16, 4608},
      {"K-major SWIZZLE_32B (SBO 256), N=48", 48, 0, 0, 6, 16, 256, 16, 256, 4096, 4608},
      {"K-major SWIZZLE_64B (SBO 512), N=48", 48, 0, 0, 4, 16, 512, 16, 512, 8192, 4608},
      {"K-major SWIZZLE_128B (SBO 1024), N=48", 48, 0, 0, 2, 16, 1024, 16, 1024, 32, 32},
      {"K-major SWIZZLE_128B (SBO 2048: 16-pixel haloed rows), N=48", 48, 0, 0, 2, 16, 2048, 16, 1024, 128, 32},
      {"K-major SWIZZLE_128B (SBO 1024), N=160", 160, 0, 0, 2, 16, 1024, 16, 1024, 32, 32},
      {"K-major SWIZZLE_128B (SBO 1024), N=256", 256, 0, 0, 2, 16, 1024, 16, 1024, 32, 32},
      {"MN-major no-swizzle, wgrad tiles (LBO 128, SBO 2304 / 2048), N=48", 48, 1, 1, 0, 128, 2304, 128, 2048, 256, 256},
      {"MN-major no-swizzle, wgrad tiles, N=80", 80, 1, 1, 0, 128, 2304, 128, 2048, 256, 256},
      {"MN-major SWIZZLE_128B (LBO 8192?, SBO 1024), N=48", 48, 1, 1, 2, 8192, 1024, 8192, 1024, 2048, 2048},
  };
  for (const Cfg& c : cfgs) {
    long long h[3] = {0, 0, 0};
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<<<1, 128, 160 * 1024>>>(c.N, c.am, c.bm, c.layout, c.albo, c.asbo, c.blbo, c.bsbo, c.astep, c.bstep, 512, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    }
    printf("%-78s issue %6.1f clk/MMA, complete %6.1f clk/MMA (floor %d)%s\n", c.name, h[0] / 512.0, h[1] / 512.0, 128 * c.N / 256, h[2] ? "" : "  TIMEOUT");
  }
  return 0;
}
