// Warp-level building blocks shared by the whole-frame kernels (critic_fused.cu, infer_fused.cu): TF32 mma.sync / ldmatrix
// wrappers, the row-sliding implicit-GEMM convolution loop and the pooling epilogue on accumulator fragments.
#pragma once
#include "common.cuh"

namespace cgs {
namespace cf {

__device__ __forceinline__ uint32_t f2tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ float tf32r(float f) { return __uint_as_float(f2tf32(f)); }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// Output rows [0, R) of one 16-pixel column strip (R even).  Iteration i loads the NK A fragments of haloed input row i
// once and feeds the three output rows i, i-1, i-2 (filter rows 0, 1, 2); epi(e, top, bot) gets the finished rows e, e+1.
template <int R, int NK, class LoadA, class Epi>
__device__ __forceinline__ void slide_rows(const float2 (&w)[3][NK], LoadA&& loadA, Epi&& epi) {
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < R + 2; ++i) {
    uint32_t a[NK][4];
    loadA(i, a);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oi = i - ky;
      if (oi >= 0 && oi < R) {
        if (ky == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[oi & 3][q] = 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < NK; ++kk)
          mma_tf32(acc[oi & 3], a[kk], __float_as_uint(w[ky][kk].x), __float_as_uint(w[ky][kk].y));
      }
    }
    if (i >= 3 && ((i - 3) & 1) == 0) epi(i - 3, acc[(i - 3) & 3], acc[(i - 2) & 3]);
  }
}

// bias + ReLU + 2x2 max-pool (first max wins, ATen's rule) on two finished rows of a strip.  The two lanes of an
// x-pair (g, g^1) split the work: even g finishes channel 2t, odd g channel 2t+1, for both pixel halves (g, g+8).
// st(h, value, idx): pooled pixel (x0 + g + 8h) >> 1, channel 2t + (g & 1); idx 4 = no gradient (ReLU off).
template <class Store>
__device__ __forceinline__ void pool2x2(const float (&top)[4], const float (&bot)[4], float bias0, float bias1, int odd, Store&& st) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float t0 = top[2 * h] + bias0, t1 = top[2 * h + 1] + bias1, b0 = bot[2 * h] + bias0, b1 = bot[2 * h + 1] + bias1;
    const float rt = __shfl_xor_sync(0xffffffffu, odd ? t0 : t1, 4), rb = __shfl_xor_sync(0xffffffffu, odd ? b0 : b1, 4);
    const float p0 = odd ? rt : t0, p1 = odd ? t1 : rt, p2 = odd ? rb : b0, p3 = odd ? b1 : rb;
    const float m01 = fmaxf(p0, p1), m23 = fmaxf(p2, p3);
    const int i01 = p1 > p0 ? 1 : 0, i23 = p3 > p2 ? 3 : 2;
    float m = fmaxf(m01, m23);
    int idx = m23 > m01 ? i23 : i01;
    if (!(m > 0.f)) { m = 0.f; idx = 4; }
    st(h, m, idx);
  }
}

// uint8 -> TF32(b / 255): 0x4B000000 | b is the float 2^23 + b; one FMA with the exactly representable constant -2^23/255
// leaves round(b * (1/255)), whose TF32 rounding equals that of the reference's b / 255.0f for all 256 values.
__device__ __forceinline__ float u8_to_tf32_unit(uint32_t b) {
  constexpr float k = 1.f / 255.f;
  return tf32r(fmaf(__uint_as_float(0x4B000000u | b), k, -8388608.f * k));
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

}  // namespace cf
}  // namespace cgs
