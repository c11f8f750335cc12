// Microbenchmark: issue rate of legacy mma.sync on sm_100a (per SM), TF32 m16n8k8 vs BF16 m16n8k16, and ldmatrix.x4.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int KIND, int CHAINS>
__global__ void k(float* out, int iters, long long* clk) {
  float c[CHAINS][4];
  for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
  for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__global__ void kld(float* out, int iters, long long* clk) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t r0, r1, r2, r3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + i * 16));
      acc += r0 ^ r1 ^ r2 ^ r3;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int KIND, int CHAINS>
void run(const char* name, int warps, int macs) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  int iters = 2000;
  k<KIND, CHAINS><<<148, warps * 32>>>(out, iters, clk);
  k<KIND, CHAINS><<<148, warps * 32>>>(out, iters, clk);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * CHAINS * warps);   // clk per mma per SM
  printf("%-28s warps %2d chains %d: %.2f clk/mma/SM  (%.0f MAC/clk/SM; latency-ish %.1f clk/mma/warp)\n", name, warps, CHAINS, per, macs / per,
         (double)h / (iters * CHAINS));
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<0, 1>("tf32 m16n8k8", 4, 1024); run<0, 4>("tf32 m16n8k8", 4, 1024); run<0, 4>("tf32 m16n8k8", 8, 1024); run<0, 4>("tf32 m16n8k8", 16, 1024);
  run<0, 8>("tf32 m16n8k8", 16, 1024);
  run<2, 4>("tf32 m16n8k4", 16, 512);
  run<1, 1>("bf16 m16n8k16", 4, 2048); run<1, 4>("bf16 m16n8k16", 4, 2048); run<1, 4>("bf16 m16n8k16", 8, 2048); run<1, 4>("bf16 m16n8k16", 16, 2048);
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  for (int w : {4, 8, 16}) {
    kld<<<148, w * 32, 32768>>>(out, 2000, clk); kld<<<148, w * 32, 32768>>>(out, 2000, clk);
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("ldmatrix.x4 warps %2d: %.2f clk per ldmatrix per SM\n", w, (double)h / (2000.0 * 8 * w));
  }
  return 0;
}
