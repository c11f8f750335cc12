// Device bodies of the fused NewCritic tail (features[9..15] + crit), shared by tail.cu and post.cu.
#pragma once
#include "head_common.cuh"

namespace cgs {


struct TailDims {
  int B, C2, C3, NB;
};

// smem carve-up shared by both kernels (floats)
struct TailSmem {
  int x2, w3, x, h, v, dh, dv, dl, dy, total;
};
__host__ __device__ inline TailSmem tail_smem(int FPC, int C2, int C3, int NB, bool bwd) {
  TailSmem s;
  int o = 0;
  s.x2 = o; o += FPC * 104 * C2;           // haloed 10x10 e2*mask: fwd pixel-major [f][slot][ci], bwd planar [f][ci][104]
  s.w3 = o; o += C2 * 9 * C3;              // conv weights as [ci][tap][co]
  s.x = o;  o += FPC * 17 * C3;            // e3*mask in the 4x4 conv's own K order (ci*17 + s)
  s.h = o;  o += FPC * NB;
  s.v = o;  o += FPC * NB;
  s.dh = o; s.dv = o; s.dl = o; s.dy = o;
  if (bwd) {
    s.dh = o; o += FPC * NB;
    s.dv = o; o += FPC * NB;
    s.dl = o; o += ((FPC + 3) & ~3);
    s.dy = o; o += FPC * 104 * C3;         // haloed 10x10 gradient of the conv output, planar [f][co][104]
  }
  s.total = o;
  return s;
}

// Stage e2*mask (haloed, zero border) and the conv weights.  PLANAR: [f][ci][slot] (backward: the wgrad loop walks
// pixels with the channel fixed) instead of pixel-major [f][slot][ci] (forward: lanes = output channels broadcast a pixel).
constexpr int TAIL_PL = 104;   // plane pitch (100 slots + 4: planes 8 banks apart)
template <int FPC, bool PLANAR>
__device__ __forceinline__ void tail_stage_in(const TailDims d, int n0, const float* __restrict__ e2, const float* __restrict__ m_e2,
                                              const float* __restrict__ w3, float* s_x2, float* s_w3) {
  const int C2 = d.C2, C3 = d.C3;
  for (int e = threadIdx.x; e < FPC * 100 * C2; e += HT) {
    const int ci = e % C2, r = e / C2;
    const int slot = r % 100, f = r / 100;
    const int yy = slot / 10, xx = slot - yy * 10;
    float v = 0.f;
    if (yy >= 1 && yy <= 8 && xx >= 1 && xx <= 8 && n0 + f < d.B) {
      const size_t o = ((size_t)(n0 + f) * 64 + (yy - 1) * 8 + (xx - 1)) * C2 + ci;
      v = __ldg(e2 + o);
      if (m_e2) v *= __ldg(m_e2 + o);
    }
    if (PLANAR) s_x2[(f * C2 + ci) * TAIL_PL + slot] = v;
    else s_x2[e] = v;
  }
  for (int e = threadIdx.x; e < C2 * 9 * C3; e += HT) {
    const int co = e % C3, r = e / C3;
    const int t = r % 9, ci = r / 9;
    s_w3[e] = __ldg(w3 + ((size_t)co * C2 + ci) * 9 + t);
  }
}

// `staged`: s_x2 (pixel-major, dropout already applied) and s_w3 are already in shared memory (fused callers).
template <int FPC>
__device__ __forceinline__ void tail_fwd_body(float* smem, const int n0, const bool staged, const TailDims d,
                                              const float* __restrict__ e2, const float* __restrict__ m_e2,
                                                      const float* __restrict__ m_e3, const float* __restrict__ m_v,
                                                      const float* __restrict__ w3, const float* __restrict__ b3,
                                                      const float* __restrict__ w14, const float* __restrict__ b14,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      float* __restrict__ e3, uint8_t* __restrict__ idx3, float* __restrict__ e4,
                                                      float* __restrict__ vout, float* __restrict__ pred) {
  const int C2 = d.C2, C3 = d.C3, NB = d.NB, B = d.B;
  const TailSmem L = tail_smem(FPC, C2, C3, NB, false);
  float *s_x2 = smem + L.x2, *s_w3 = smem + L.w3, *s_x = smem + L.x, *s_h = smem + L.h, *s_v = smem + L.v;
  const int K = 16 * C3, ldx = 17 * C3;
  if (!staged) tail_stage_in<FPC, false>(d, n0, e2, m_e2, w3, s_x2, s_w3);
  __syncthreads();

  // ---- features.10-12: conv + ReLU + 2x2 max-pool; one (frame, window, co) item per thread iteration
  for (int it = threadIdx.x; it < FPC * 16 * C3; it += HT) {
    const int co = it % C3, r = it / C3;
    const int win = r & 15, f = r >> 4;
    const int wy = win >> 2, wx = win & 3;
    const float* xb = s_x2 + ((size_t)f * 100 + (2 * wy) * 10 + 2 * wx) * C2;   // top-left of the 4x4 input patch
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int ci = 0; ci < C2; ++ci) {
      float in[4][4];
#pragma unroll
      for (int yy = 0; yy < 4; ++yy)
#pragma unroll
        for (int xx = 0; xx < 4; ++xx) in[yy][xx] = xb[(yy * 10 + xx) * C2 + ci];
      const float* wp = s_w3 + (size_t)ci * 9 * C3 + co;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float w = wp[t * C3];
        const int ky = t / 3, kx = t % 3;
        a0 = fmaf(in[ky][kx], w, a0);
        a1 = fmaf(in[ky][kx + 1], w, a1);
        a2 = fmaf(in[ky + 1][kx], w, a2);
        a3 = fmaf(in[ky + 1][kx + 1], w, a3);
      }
    }
    const float bb = __ldg(b3 + co);
    float m = fmaxf(a0 + bb, 0.f);
    int am = 0;
    float t1 = fmaxf(a1 + bb, 0.f); if (t1 > m) { m = t1; am = 1; }
    float t2 = fmaxf(a2 + bb, 0.f); if (t2 > m) { m = t2; am = 2; }
    float t3 = fmaxf(a3 + bb, 0.f); if (t3 > m) { m = t3; am = 3; }
    float xm = 0.f;
    if (n0 + f < B) {
      const size_t o = ((size_t)(n0 + f) * 16 + win) * C3 + co;
      e3[o] = m;
      idx3[o] = (uint8_t)am;
      xm = m_e3 ? m * __ldg(m_e3 + o) : m;
    }
    s_x[f * ldx + co * 17 + win] = xm;
  }
  __syncthreads();

  // ---- features.14-15: h = relu(W14 x + b14), dot products in the filter's own K order
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
  for (int e = threadIdx.x; e < FPC * NB; e += HT) {
    const int f = e / NB;
    if (n0 + f < B) {
      const size_t o = (size_t)(n0 + f) * NB + (e - f * NB);
      vout[o] = s_v[e];
      if (m_v) s_v[e] *= __ldg(m_v + o);
    }
  }
  __syncthreads();
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

struct TailGrads {
  float *dw3, *db3, *dw14, *db14, *dw1, *db1, *dw2, *db2;
};

// de2 (global) and/or s_de2 (shared, [f][64][C2], for fused callers) receive the gradient w.r.t. e2.
template <int FPC>
__device__ __forceinline__ void tail_bwd_body(float* smem, const int n0, const TailDims d, const float* __restrict__ e2,
                                              const float* __restrict__ m_e2,
                                                      const float* __restrict__ m_e3, const float* __restrict__ m_v,
                                                      const float* __restrict__ w3, const float* __restrict__ w14,
                                                      const float* __restrict__ w1, const float* __restrict__ w2,
                                                      const float* __restrict__ e3, const uint8_t* __restrict__ idx3,
                                                      const float* __restrict__ e4, const float* __restrict__ v,
                                                      const float* __restrict__ pred, const float* __restrict__ dpred,
                                                      const float* __restrict__ de3_ext, const float* __restrict__ de4,
                                                      const TailGrads G, float* __restrict__ de2, float* s_de2) {
  const int C2 = d.C2, C3 = d.C3, NB = d.NB, B = d.B;
  const TailSmem L = tail_smem(FPC, C2, C3, NB, true);
  float *s_x2 = smem + L.x2, *s_w3 = smem + L.w3, *s_x = smem + L.x, *s_h = smem + L.h, *s_vd = smem + L.v;
  float *s_dh = smem + L.dh, *s_dv = smem + L.dv, *s_dl = smem + L.dl, *s_dy = smem + L.dy;
  const int K = 16 * C3, ldx = 17 * C3;
  const bool wg = G.dw14 != nullptr;

  tail_stage_in<FPC, true>(d, n0, e2, m_e2, w3, s_x2, s_w3);
  for (int e = threadIdx.x; e < FPC * TAIL_PL * C3; e += HT) s_dy[e] = 0.f;
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
      const float dl = (dpred ? __ldg(dpred + n0 + f) : 0.f) * pr * (1.f - pr);
      dv = dl * __ldg(w2 + (e - f * NB));
      if (m_v) { const float mk = __ldg(m_v + o); dv *= mk; vv *= mk; }
      if (!(__ldg(v + o) > 0.f)) dv = 0.f;
    }
    s_h[e] = hv; s_vd[e] = vv; s_dv[e] = dv;
  }
  if (threadIdx.x < FPC) {
    const int f = threadIdx.x;
    float dl = 0.f;
    if (n0 + f < B && dpred) { const float pr = __ldg(pred + n0 + f); dl = __ldg(dpred + n0 + f) * pr * (1.f - pr); }
    s_dl[f] = dl;
  }
  __syncthreads();
  if (wg) {
    dense_dw<FPC>(s_dl, 1, s_vd, NB, NB, 1, G.dw2, G.db2);
    dense_dw<FPC>(s_dv, NB, s_h, NB, NB, NB, G.dw1, G.db1);
  }
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
    for (int e = threadIdx.x; e < NB * K; e += HT) {
      const int n = e / K, k = e - n * K;
      const int sl = kw_slot(k);
      float acc = 0.f;
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc = fmaf(s_dh[f * NB + n], s_x[f * ldx + sl], acc);
      atomicAdd(G.dw14 + e, acc);
    }
    for (int n = threadIdx.x; n < NB; n += HT) {
      float acc = 0.f;
#pragma unroll
      for (int f = 0; f < FPC; ++f) acc += s_dh[f * NB + n];
      atomicAdd(G.db14 + n, acc);
    }
  }
  __syncthreads();   // all reads of s_x done
  // ---- gradient reaching embeds[3] from the head (+ from the decoder), through dropout, ReLU and the pool:
  //      scattered to the arg-max position of each window in the haloed s_dy
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
    const int co = k >> 4, win = k & 15;         // filter K order: k = ci*16 + s
    const int wy = win >> 2, wx = win & 3;
#pragma unroll
    for (int f = 0; f < FPC; ++f) {
      if (n0 + f >= B) continue;
      const size_t o = ((size_t)(n0 + f) * 16 + win) * C3 + co;
      float g = acc[f];
      if (m_e3) g *= __ldg(m_e3 + o);
      if (de3_ext) g += __ldg(de3_ext + o);
      if (!(__ldg(e3 + o) > 0.f)) g = 0.f;
      const int pos = idx3[o];
      s_dy[(f * C3 + co) * TAIL_PL + (2 * wy + (pos >> 1) + 1) * 10 + (2 * wx + (pos & 1) + 1)] = g;
    }
  }
  __syncthreads();
  // ---- features.10 wgrad / bias grad: dW3[co][ci][t] += sum_{f,y,x} dY[f,y,x,co] * X[f,y+ky-1,x+kx-1,ci]
  if (G.dw3) {
    for (int e = threadIdx.x; e < C3 * C2 * 9; e += HT) {
      const int t = e % 9, r = e / 9;
      const int ci = r % C2, co = r / C2;
      const int ky = t / 3, kx = t - ky * 3;
      float acc = 0.f;
      for (int f = 0; f < FPC; ++f) {
        const float* dyb = s_dy + (f * C3 + co) * TAIL_PL + 11;
        const float* xb = s_x2 + (f * C2 + ci) * TAIL_PL + ky * 10 + kx;
#pragma unroll 2
        for (int y = 0; y < 8; ++y)
#pragma unroll
          for (int x = 0; x < 8; ++x) acc = fmaf(dyb[y * 10 + x], xb[y * 10 + x], acc);
      }
      atomicAdd(G.dw3 + e, acc);
    }
    for (int co = threadIdx.x; co < C3; co += HT) {
      float acc = 0.f;
      for (int f = 0; f < FPC; ++f)
        for (int s = 0; s < 100; ++s) acc += s_dy[(f * C3 + co) * TAIL_PL + s];
      atomicAdd(G.db3 + co, acc);
    }
  }
  // ---- features.10 dgrad through the input dropout: de2[f,y,x,ci] = m * sum_{t,co} dY[f,y-ky+1,x-kx+1,co] * W3[co][ci][t]
  if (de2 || s_de2) {
    for (int e = threadIdx.x; e < FPC * 64 * C2; e += HT) {
      const int ci = e % C2, r = e / C2;
      const int px = r & 63, f = r >> 6;
      const int y = px >> 3, x = px & 7;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int ky = t / 3, kx = t - ky * 3;
        const float* dyb = s_dy + (f * C3) * TAIL_PL + (y + 2 - ky) * 10 + (x + 2 - kx);
        const float* wp = s_w3 + ((size_t)ci * 9 + t) * C3;
#pragma unroll 4
        for (int co = 0; co < C3; ++co) acc = fmaf(dyb[co * TAIL_PL], wp[co], acc);
      }
      if (n0 + f < B) {
        const size_t o = ((size_t)(n0 + f) * 64 + px) * C2 + ci;
        if (m_e2) acc *= __ldg(m_e2 + o);
        if (de2) de2[o] = acc;
      } else {
        acc = 0.f;
      }
      if (s_de2) s_de2[e] = acc;
    }
  }
}


}  // namespace cgs
