// Building blocks shared by head.cu (NewCritic tail, unfused) and tail.cu (features.10 .. crit fused).
#pragma once
#include "common.cuh"

namespace cgs {

constexpr int HT = 256;       // threads per CTA

// out[f][n] = sum_k in[f][k] * w[n][k] (+bias) ; warp per n, lanes along k.  `in` in smem
// with row stride ld_in, `w` in global.  act: 0 none, 1 relu, 2 sigmoid.
template <int FPC, int ACT>
__device__ void dense_rows(const float* s_in, int ld_in, const float* __restrict__ w, const float* __restrict__ bias,
                           int K, int N, float* s_out, int ld_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = HT / 32;
  for (int n = warp; n < N; n += nw) {
    float acc[FPC];
#pragma unroll
    for (int f = 0; f < FPC; ++f) acc[f] = 0.f;
    const float* wr = w + (size_t)n * K;
#pragma unroll 4
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc[f] = fmaf(s_in[f * ld_in + k], wv, acc[f]);
    }
#pragma unroll
    for (int f = 0; f < FPC; ++f) acc[f] = warp_sum(acc[f]);
    if (lane < FPC) {
      float v = 0.f;
#pragma unroll
      for (int f = 0; f < FPC; ++f) if (lane == f) v = acc[f];
      v += bias ? __ldg(bias + n) : 0.f;
      if (ACT == 1) v = fmaxf(v, 0.f);
      if (ACT == 2) v = 1.f / (1.f + expf(-v));
      s_out[lane * ld_out + n] = v;
    }
  }
}

// din[f][k] = sum_n dout[f][n] * w[n][k] ; thread per k (coalesced w rows).
template <int FPC>
__device__ void dense_din(const float* s_dout, int ld_do, const float* __restrict__ w, int K, int N,
                          float* s_din, int ld_di) {
  for (int k = threadIdx.x; k < K; k += HT) {
    float acc[FPC];
#pragma unroll
    for (int f = 0; f < FPC; ++f) acc[f] = 0.f;
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
      const float wv = __ldg(w + (size_t)n * K + k);
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc[f] = fmaf(s_dout[f * ld_do + n], wv, acc[f]);
    }
#pragma unroll
    for (int f = 0; f < FPC; ++f) s_din[f * ld_di + k] = acc[f];
  }
}

// dw[n][k] += sum_f dout[f][n] * in[f][k] ; db[n] += sum_f dout[f][n].
template <int FPC>
__device__ void dense_dw(const float* s_dout, int ld_do, const float* s_in, int ld_in, int K, int N,
                         float* __restrict__ dw, float* __restrict__ db) {
  for (int e = threadIdx.x; e < N * K; e += HT) {
    const int n = e / K, k = e - n * K;
    float acc = 0.f;
#pragma unroll
    for (int f = 0; f < FPC; ++f) acc = fmaf(s_dout[f * ld_do + n], s_in[f * ld_in + k], acc);
    atomicAdd(dw + e, acc);
  }
  if (db)
    for (int n = threadIdx.x; n < N; n += HT) {
      float acc = 0.f;
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc += s_dout[f * ld_do + n];
      atomicAdd(db + n, acc);
    }
}

// e3 is NHWC [B,4,4,C3]: element (s = y*4+x, ci).  The OIHW filter w14[n][ci][s] is read
// coalesced along its own K order kw = ci*16 + s, so the frame is staged in smem in that
// order with a 17-float pitch per channel (conflict-free both for staging and the dots).
__device__ __forceinline__ int kw_slot(int kw) { return (kw >> 4) * 17 + (kw & 15); }


}  // namespace cgs
