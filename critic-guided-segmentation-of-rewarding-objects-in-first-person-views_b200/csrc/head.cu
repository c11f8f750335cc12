// NewCritic tail (features[13..15] + crit, reference nets.py:183-195) and the decoder's
// bottleneck 1x1 conv (dec[4], nets.py:484) as fused dense kernels.
//
// A CTA owns FPC frames; activations of those frames live in shared memory for the whole
// chain  e3 --(4x4 valid conv == dense K=16*C3)--> e4 --Linear--> v --Linear--> pred,
// weights stream from L2 with coalesced reads (lanes along K for the forward dot
// products, lanes along K for dIn / dW in backward).  Parameter gradients are reduced
// over the CTA's frames in registers and pushed with one RED per weight per CTA.
#include <type_traits>
#include "head_common.cuh"

namespace cgs {

// Head forward kernel.  smem: x [FPC][ldx], h [FPC][NB], v [FPC][NB].
template <int FPC>
__global__ void __launch_bounds__(HT) head_fwd_kernel(const float* __restrict__ e3, const float* __restrict__ m_e3,
                                                      const float* __restrict__ m_v, const float* __restrict__ w14,
                                                      const float* __restrict__ b14, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ w2,
                                                      const float* __restrict__ b2, int B, int C3, int NB,
                                                      float* __restrict__ e4, float* __restrict__ vout,
                                                      float* __restrict__ pred) {
  extern __shared__ __align__(16) float smem[];
  const int K = 16 * C3, ldx = 17 * C3;
  float* s_x = smem;
  float* s_h = s_x + FPC * ldx;
  float* s_v = s_h + FPC * NB;
  const int n0 = blockIdx.x * FPC;
  // stage x = e3 * mask, remapped NHWC (s*C3+ci) -> kw (ci*16+s)
  for (int e = threadIdx.x; e < FPC * K; e += HT) {
    const int f = e / K, k = e - f * K;
    const int s = k / C3, ci = k - s * C3;
    float x = 0.f;
    if (n0 + f < B) {
      const size_t o = (size_t)(n0 + f) * K + k;
      x = __ldg(e3 + o);
      if (m_e3) x *= __ldg(m_e3 + o);
    }
    s_x[f * ldx + ci * 17 + s] = x;
  }
  __syncthreads();
  // h = relu(W14 x + b14): dot products in kw order, skipping the pad slot via kw_slot
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = HT / 32;
    for (int n = warp; n < NB; n += nw) {
      float acc[FPC];
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc[f] = 0.f;
      const float* wr = w14 + (size_t)n * K;
#pragma unroll 4
      for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(wr + k);
        const int sl = kw_slot(k);
#pragma unroll
        for (int f = 0; f < FPC; ++f) acc[f] = fmaf(s_x[f * ldx + sl], wv, acc[f]);
      }
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc[f] = warp_sum(acc[f]);
      if (lane < FPC) {
        float v = 0.f;
#pragma unroll
        for (int f = 0; f < FPC; ++f) if (lane == f) v = acc[f];
        v = fmaxf(v + __ldg(b14 + n), 0.f);
        s_h[lane * NB + n] = v;
        if (n0 + lane < B) e4[(size_t)(n0 + lane) * NB + n] = v;
      }
    }
  }
  __syncthreads();
  dense_rows<FPC, 1>(s_h, NB, w1, b1, NB, NB, s_v, NB);
  __syncthreads();
  // save v (pre-dropout), apply dropout in place
  for (int e = threadIdx.x; e < FPC * NB; e += HT) {
    const int f = e / NB;
    if (n0 + f < B) {
      const size_t o = (size_t)(n0 + f) * NB + (e - f * NB);
      vout[o] = s_v[e];
      if (m_v) s_v[e] *= __ldg(m_v + o);
    }
  }
  __syncthreads();
  // pred = sigmoid(w2 . v + b2): one warp per frame
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = HT / 32;
    for (int f = warp; f < FPC; f += nw) {
      float acc = 0.f;
      for (int k = lane; k < NB; k += 32) acc = fmaf(s_v[f * NB + k], __ldg(w2 + k), acc);
      acc = warp_sum(acc);
      if (lane == 0 && n0 + f < B) pred[n0 + f] = 1.f / (1.f + expf(-(acc + __ldg(b2))));
    }
  }
}

// Head backward kernel.  smem: x [FPC][ldx] (later reused for dx), h, v(dropped), dh, dv: [FPC][NB] each, dl [FPC].
template <int FPC>
__global__ void __launch_bounds__(HT) head_bwd_kernel(const float* __restrict__ e3, const float* __restrict__ m_e3,
                                                      const float* __restrict__ m_v, const float* __restrict__ w14,
                                                      const float* __restrict__ w1, const float* __restrict__ w2,
                                                      const float* __restrict__ e4, const float* __restrict__ v,
                                                      const float* __restrict__ pred, const float* __restrict__ dpred,
                                                      const float* __restrict__ de4, int B, int C3, int NB,
                                                      float* dw14, float* db14, float* dw1, float* db1, float* dw2,
                                                      float* db2, float* __restrict__ de3) {
  extern __shared__ __align__(16) float smem[];
  const int K = 16 * C3, ldx = 17 * C3;
  float* s_x = smem;
  float* s_h = s_x + FPC * ldx;
  float* s_vd = s_h + FPC * NB;   // v after dropout (input of crit.4)
  float* s_dh = s_vd + FPC * NB;  // grad wrt pre-activation of features.14
  float* s_dv = s_dh + FPC * NB;  // grad wrt pre-activation of crit.1
  float* s_dl = s_dv + FPC * NB;  // grad wrt logit
  const int n0 = blockIdx.x * FPC;
  const bool wg = dw14 != nullptr;

  for (int e = threadIdx.x; e < FPC * K; e += HT) {
    const int f = e / K, k = e - f * K;
    const int s = k / C3, ci = k - s * C3;
    float x = 0.f;
    if (n0 + f < B) {
      const size_t o = (size_t)(n0 + f) * K + k;
      x = __ldg(e3 + o);
      if (m_e3) x *= __ldg(m_e3 + o);
    }
    s_x[f * ldx + ci * 17 + s] = x;
  }
  for (int e = threadIdx.x; e < FPC * NB; e += HT) {
    const int f = e / NB;
    float hv = 0.f, vv = 0.f, dv = 0.f;
    if (n0 + f < B) {
      const size_t o = (size_t)(n0 + f) * NB + (e - f * NB);
      hv = __ldg(e4 + o);
      vv = __ldg(v + o);
      const float pr = __ldg(pred + n0 + f);
      const float dl = __ldg(dpred + n0 + f) * pr * (1.f - pr);
      // d v_dropped = dl * w2 ; through dropout and ReLU of crit.1
      dv = dl * __ldg(w2 + (e - f * NB));
      if (m_v) { const float mk = __ldg(m_v + o); dv *= mk; vv *= mk; }
      if (!(__ldg(v + o) > 0.f)) dv = 0.f;
    }
    s_h[e] = hv; s_vd[e] = vv; s_dv[e] = dv;
  }
  if (threadIdx.x < FPC) {
    const int f = threadIdx.x;
    float dl = 0.f;
    if (n0 + f < B) { const float pr = __ldg(pred + n0 + f); dl = __ldg(dpred + n0 + f) * pr * (1.f - pr); }
    s_dl[f] = dl;
  }
  __syncthreads();
  if (wg) {
    // crit.4: dw2[k] += sum_f dl[f]*vd[f][k], db2 += sum_f dl
    dense_dw<FPC>(s_dl, 1, s_vd, NB, NB, 1, dw2, db2);
    // crit.1: dw1[n][k] += sum_f dv[f][n]*h[f][k]
    dense_dw<FPC>(s_dv, NB, s_h, NB, NB, NB, dw1, db1);
  }
  // dh = (dv W1 + de4) * (h > 0)
  dense_din<FPC>(s_dv, NB, w1, NB, NB, s_dh, NB);
  __syncthreads();
  for (int e = threadIdx.x; e < FPC * NB; e += HT) {
    const int f = e / NB;
    float g = s_dh[e];
    if (de4 && n0 + f < B) g += __ldg(de4 + (size_t)(n0 + f) * NB + (e - f * NB));
    s_dh[e] = (s_h[e] > 0.f) ? g : 0.f;
  }
  __syncthreads();
  if (wg) {
    // features.14: dw14[n][kw] += sum_f dh[f][n] * x[f][slot(kw)]
    for (int e = threadIdx.x; e < NB * K; e += HT) {
      const int n = e / K, k = e - n * K;
      const int sl = kw_slot(k);
      float acc = 0.f;
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc = fmaf(s_dh[f * NB + n], s_x[f * ldx + sl], acc);
      atomicAdd(dw14 + e, acc);
    }
    for (int n = threadIdx.x; n < NB; n += HT) {
      float acc = 0.f;
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc += s_dh[f * NB + n];
      atomicAdd(db14 + n, acc);
    }
  }
  if (de3 == nullptr) return;
  __syncthreads();   // all reads of s_x done; reuse it for dx in kw-slot order
  for (int k = threadIdx.x; k < K; k += HT) {
    float acc[FPC];
#pragma unroll
    for (int f = 0; f < FPC; ++f) acc[f] = 0.f;
#pragma unroll 4
    for (int n = 0; n < NB; ++n) {
      const float wv = __ldg(w14 + (size_t)n * K + k);
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc[f] = fmaf(s_dh[f * NB + n], wv, acc[f]);
    }
    const int sl = kw_slot(k);
#pragma unroll
    for (int f = 0; f < FPC; ++f) s_x[f * ldx + sl] = acc[f];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < FPC * K; e += HT) {
    const int f = e / K, k = e - f * K;
    const int s = k / C3, ci = k - s * C3;
    if (n0 + f < B) {
      const size_t o = (size_t)(n0 + f) * K + k;
      float g = s_x[f * ldx + ci * 17 + s];
      if (m_e3) g *= __ldg(m_e3 + o);
      de3[o] = g;
    }
  }
}

// Generic dense forward: out[B,N] = in[B,K] w[N,K]^T + bias.
template <int FPC>
__global__ void __launch_bounds__(HT) dense_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int B, int K, int N,
                                                       float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;            // [FPC][K]
  float* s_out = smem + FPC * K; // [FPC][N]
  const int n0 = blockIdx.x * FPC;
  for (int e = threadIdx.x; e < FPC * K; e += HT) {
    const int f = e / K;
    s_in[e] = (n0 + f < B) ? __ldg(in + (size_t)n0 * K + e) : 0.f;
  }
  __syncthreads();
  dense_rows<FPC, 0>(s_in, K, w, bias, K, N, s_out, N);
  __syncthreads();
  for (int e = threadIdx.x; e < FPC * N; e += HT) {
    const int f = e / N;
    if (n0 + f < B) out[(size_t)n0 * N + e] = s_out[e];
  }
}

template <int FPC>
__global__ void __launch_bounds__(HT) dense_bwd_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ dout, int B, int K, int N,
                                                       float* __restrict__ din, float* dw, float* db) {
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;              // [FPC][K]
  float* s_do = s_in + FPC * K;    // [FPC][N]
  float* s_di = s_do + FPC * N;    // [FPC][K]
  const int n0 = blockIdx.x * FPC;
  for (int e = threadIdx.x; e < FPC * K; e += HT) {
    const int f = e / K;
    s_in[e] = (n0 + f < B) ? __ldg(in + (size_t)n0 * K + e) : 0.f;
  }
  for (int e = threadIdx.x; e < FPC * N; e += HT) {
    const int f = e / N;
    s_do[e] = (n0 + f < B) ? __ldg(dout + (size_t)n0 * N + e) : 0.f;
  }
  __syncthreads();
  if (dw) dense_dw<FPC>(s_do, N, s_in, K, K, N, dw, db);
  if (din) {
    dense_din<FPC>(s_do, N, w, K, N, s_di, K);
    __syncthreads();
    for (int e = threadIdx.x; e < FPC * K; e += HT) {
      const int f = e / K;
      if (n0 + f < B) din[(size_t)n0 * K + e] = s_di[e];
    }
  }
}

// Few frames per CTA while the batch is small (the chain is latency-bound: more CTAs in flight),
// 16 once there are enough frames to fill the SMs (weights are re-read from L2 once per CTA).
static int frames_per_cta(int B) { return B <= 2048 ? 4 : 16; }

static int set_smem(const void* fn, size_t bytes) {
  if (bytes > 227 * 1024) { set_error("dense/head: shared memory request %zu exceeds 227 KB", bytes); return CGS_EUNSUPPORTED; }
  if (bytes > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return 0;
}

}  // namespace cgs

extern "C" int cgs_head_fwd(const float* e3, const float* m_e3, const float* m_v, const float* w14, const float* b14,
                            const float* w1, const float* b1, const float* w2, const float* b2, int32_t B, int32_t C3,
                            int32_t NB, float* e4, float* v, float* pred, void* stream) {
  using namespace cgs;
  CGS_REQUIRE(e3 && w14 && b14 && w1 && b1 && w2 && b2 && e4 && v && pred, "head_fwd: null pointer");
  CGS_REQUIRE(B > 0 && C3 > 0 && NB > 0, "head_fwd: bad sizes");
  auto go = [&](auto fpc_tag) {
    constexpr int FPC = decltype(fpc_tag)::value;
    size_t smem = ((size_t)FPC * 17 * C3 + 2 * FPC * NB) * sizeof(float);
    if (int e = set_smem((const void*)head_fwd_kernel<FPC>, smem)) return e;
    head_fwd_kernel<FPC><<<(B + FPC - 1) / FPC, HT, smem, (cudaStream_t)stream>>>(e3, m_e3, m_v, w14, b14, w1, b1, w2, b2, B,
                                                                                 C3, NB, e4, v, pred);
    return check_launch("head_fwd");
  };
  return frames_per_cta(B) == 4 ? go(std::integral_constant<int, 4>{}) : go(std::integral_constant<int, 16>{});
}

extern "C" int cgs_head_bwd(const float* e3, const float* m_e3, const float* m_v, const float* w14, const float* w1,
                            const float* w2, const float* e4, const float* v, const float* pred, const float* dpred,
                            const float* de4, int32_t B, int32_t C3, int32_t NB, float* dw14, float* db14, float* dw1,
                            float* db1, float* dw2, float* db2, float* de3, void* stream) {
  using namespace cgs;
  CGS_REQUIRE(e3 && w14 && w1 && w2 && e4 && v && pred && dpred, "head_bwd: null pointer");
  CGS_REQUIRE(B > 0 && C3 > 0 && NB > 0, "head_bwd: bad sizes");
  if (dw14) CGS_REQUIRE(db14 && dw1 && db1 && dw2 && db2, "head_bwd: parameter gradients are all-or-none");
  auto go = [&](auto fpc_tag) {
    constexpr int FPC = decltype(fpc_tag)::value;
    size_t smem = ((size_t)FPC * 17 * C3 + 4 * FPC * NB + FPC) * sizeof(float);
    if (int e = set_smem((const void*)head_bwd_kernel<FPC>, smem)) return e;
    head_bwd_kernel<FPC><<<(B + FPC - 1) / FPC, HT, smem, (cudaStream_t)stream>>>(e3, m_e3, m_v, w14, w1, w2, e4, v, pred,
                                                                                 dpred, de4, B, C3, NB, dw14, db14, dw1, db1,
                                                                                 dw2, db2, de3);
    return check_launch("head_bwd");
  };
  return frames_per_cta(B) == 4 ? go(std::integral_constant<int, 4>{}) : go(std::integral_constant<int, 16>{});
}

extern "C" int cgs_dense_fwd(const float* in, const float* w, const float* bias, int32_t B, int32_t K, int32_t N,
                             float* out, void* stream) {
  using namespace cgs;
  CGS_REQUIRE(in && w && out && B > 0 && K > 0 && N > 0, "dense_fwd: bad args");
  auto go = [&](auto fpc_tag) {
    constexpr int FPC = decltype(fpc_tag)::value;
    size_t smem = (size_t)FPC * (K + N) * sizeof(float);
    if (int e = set_smem((const void*)dense_fwd_kernel<FPC>, smem)) return e;
    dense_fwd_kernel<FPC><<<(B + FPC - 1) / FPC, HT, smem, (cudaStream_t)stream>>>(in, w, bias, B, K, N, out);
    return check_launch("dense_fwd");
  };
  return frames_per_cta(B) == 4 ? go(std::integral_constant<int, 4>{}) : go(std::integral_constant<int, 16>{});
}

extern "C" int cgs_dense_bwd(const float* in, const float* w, const float* dout, int32_t B, int32_t K, int32_t N,
                             float* din, float* dw, float* db, void* stream) {
  using namespace cgs;
  CGS_REQUIRE(in && w && dout && B > 0 && K > 0 && N > 0, "dense_bwd: bad args");
  auto go = [&](auto fpc_tag) {
    constexpr int FPC = decltype(fpc_tag)::value;
    size_t smem = (size_t)FPC * (2 * K + N) * sizeof(float);
    if (int e = set_smem((const void*)dense_bwd_kernel<FPC>, smem)) return e;
    dense_bwd_kernel<FPC><<<(B + FPC - 1) / FPC, HT, smem, (cudaStream_t)stream>>>(in, w, dout, B, K, N, din, dw, db);
    return check_launch("dense_bwd");
  };
  return frames_per_cta(B) == 4 ? go(std::integral_constant<int, 4>{}) : go(std::integral_constant<int, 16>{});
}
