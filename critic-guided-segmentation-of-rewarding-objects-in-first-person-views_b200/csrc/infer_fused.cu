// `-process` inference, encoder + decoder half (reference main.py:1139-1150, nets.py:197-212 and 494-517), as ONE persistent
// kernel for the chfak = 1 geometry: uint8 frame -> /255 -> NewCritic.forward(collect=True) (eval mode) -> UnetDecoder's
// dec[4] .. dec[0] -> o0 [B,32,32,8] (the feature map the masker convolutions consume) and pred [B].
//
// Same design as critic_fused.cu: one CTA (16 warps) owns a frame, every activation lives in shared memory as haloed
// 4-channel half-planes, convolutions are implicit GEMMs on mma.sync TF32 fed by ldmatrix.  New here:
//   * nn.Upsample(nearest) + T.cat never exist as data: the concatenated operand is two k-steps per filter tap, and the
//     upsampled half is read straight from the low-resolution plane - ldmatrix takes one row address per lane, so lane
//     addresses simply repeat (haloed virtual coordinate v -> source coordinate (v + 1) >> 1);
//   * the small-map stages dec[3] (4x4) and dec[2] (8x8) split K over the warps and reduce through shared memory;
//   * decoder weights are packed once per checkpoint into mma B-fragment order (cgs_infer_pack_decoder) and read from L2.
// The two masker convolutions (62 % of the MACs, 64x64 maps of 16 channels) stay with the tcgen05 kernel (conv_tc.cu).
#include "fused_common.cuh"

namespace cgs {
namespace inf {
using namespace cf;

constexpr int NT = 512;
constexpr int P0 = 66, SX = 66 * 66 * 4;
constexpr int P1 = 34, PL1 = 34 * 34 * 4;
constexpr int P2 = 18, PL2 = 18 * 18 * 4;
constexpr int P3 = 10, PL3 = 10 * 10 * 4;
constexpr int P4 = 6, PL4 = 6 * 6 * 4;
// shared memory map (float offsets)
constexpr int oX = 0;                               // the frame; after conv0: scratch of the split-K reductions
constexpr int oScr = 0;
constexpr int oE0 = oX + SX;
constexpr int oE1 = oE0 + 2 * PL1, oO1 = oE1 + 2 * PL2;
constexpr int oE2 = oO1 + 2 * PL2, oO2 = oE2 + 2 * PL3;
constexpr int oCat3 = oO2 + 2 * PL3;                // 12 planes @4x4: e3 (16 ch) | dec[4] output broadcast (32 ch)
constexpr int oO3 = oCat3 + 12 * PL4;               // 4 planes @4x4
constexpr int oX3 = oO3 + 4 * PL4;                  // e3 in the 4x4 conv's K order
constexpr int oHead = oX3 + 256;                    // h[32] v[32]
constexpr int oU8 = oHead + 64;
constexpr int oW = oU8 + 3072;                      // critic conv fragments
constexpr int wL0 = 0, wL1 = 384, wL2 = wL1 + 576, wL3 = wL2 + 576, szW = wL3 + 1152;
constexpr int oBias = oW + szW;                     // b0[8] b1[8] b2[8] b3[16] bd3[16] bd2[8] bd1[8] bd0[8]
constexpr int oHW = oBias + 96;                     // wl1[1024] bl1[32] wl2[32] bl2[1..] b4[32] wd4[1024] bd4[32]
constexpr int hWl1 = 0, hBl1 = 1024, hWl2 = 1056, hBl2 = 1088, hB4 = 1092, hWd4 = 1124, hBd4 = 2148, szHW = 2180;
constexpr int SMEM_FLOATS = oHW + szHW;
static_assert(SMEM_FLOATS * 4 <= 227 * 1024, "shared memory budget");
// packed decoder fragments (global): [step][lane] float2, step = (tap * NPG + pg) * NT + nt
constexpr int pkD3 = 0, pkD2 = pkD3 + 108 * 64, pkD1 = pkD2 + 27 * 64, pkD0 = pkD1 + 18 * 64, PACK_FLOATS = pkD0 + 18 * 64;

struct Params {
  const uint8_t* frames;
  const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4, *wl1, *bl1, *wl2, *bl2;
  const float *wd4, *bd4, *bd3, *bd2, *bd1, *bd0;
  const float* pack;
  float* pred;
  float* o0;
  int B;
};

__global__ void pack_decoder_kernel(const float* __restrict__ w3, const float* __restrict__ w2, const float* __restrict__ w1,
                                    const float* __restrict__ w0, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= PACK_FLOATS / 2) return;
  int s = e >> 5;
  const int lane = e & 31, g = lane >> 2, t = lane & 3;
  const float* w;
  int npg, nt_, cin;
  if (s < 108) { w = w3; npg = 6; nt_ = 2; cin = 48; }
  else if ((s -= 108) < 27) { w = w2; npg = 3; nt_ = 1; cin = 24; }
  else if ((s -= 27) < 18) { w = w1; npg = 2; nt_ = 1; cin = 16; }
  else { s -= 18; w = w0; npg = 2; nt_ = 1; cin = 16; }
  const int nt = s % nt_, pg = (s / nt_) % npg, tap = s / (nt_ * npg);
  const int co = nt * 8 + g, ci = pg * 8 + t;
  out[2 * e] = tf32r(__ldg(w + ((size_t)co * cin + ci) * 9 + tap));
  out[2 * e + 1] = tf32r(__ldg(w + ((size_t)co * cin + ci + 4) * 9 + tap));
}

__device__ __forceinline__ void stage_frame(const uint8_t* __restrict__ sU8, float* __restrict__ sXd, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = tid + NT * i, y = p >> 6, x = p & 63;
    const uint8_t* s = sU8 + p * 3;
    float4 v;
    v.x = u8_to_tf32_unit(s[0]); v.y = u8_to_tf32_unit(s[1]); v.z = u8_to_tf32_unit(s[2]); v.w = 0.f;
    *reinterpret_cast<float4*>(sXd + ((y + 1) * P0 + x + 1) * 4) = v;
  }
  if (tid < 260) {   // halo ring (the region doubles as reduction scratch, so it is rewritten every frame)
    int y, x;
    if (tid < 66) { y = 0; x = tid; }
    else if (tid < 132) { y = 65; x = tid - 66; }
    else if (tid < 196) { y = tid - 131; x = 0; }
    else { y = tid - 195; x = 65; }
    *reinterpret_cast<float4*>(sXd + (y * P0 + x) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ void prefetch_frame(const uint8_t* src, float* sm, int tid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(sm + oU8);
  for (int c = tid; c < 768; c += NT)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + c * 16), "l"(src + c * 16));
  asm volatile("cp.async.commit_group;\n" ::);
}

// bias + ReLU + 2x2 max-pool of two finished rows; value only (no arg-max: inference)
template <class Store>
__device__ __forceinline__ void pool_fwd(const float (&top)[4], const float (&bot)[4], float bias0, float bias1, int odd, Store&& st) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float t0 = top[2 * h] + bias0, t1 = top[2 * h + 1] + bias1, b0 = bot[2 * h] + bias0, b1 = bot[2 * h + 1] + bias1;
    const float rt = __shfl_xor_sync(0xffffffffu, odd ? t0 : t1, 4), rb = __shfl_xor_sync(0xffffffffu, odd ? b0 : b1, 4);
    const float own = odd ? fmaxf(t1, b1) : fmaxf(t0, b0);
    st(h, fmaxf(fmaxf(own, fmaxf(rt, rb)), 0.f));
  }
}

__global__ void __launch_bounds__(NT, 1) infer_fused_kernel(const Params p) {
  extern __shared__ __align__(128) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, odd = g & 1;
  const int lj = lane >> 3, lr = lane & 7;
  float *sH = sm + oHead, *sV = sH + 32;
  const float2* sWf = reinterpret_cast<const float2*>(sm + oW);
  const float2* pk = reinterpret_cast<const float2*>(p.pack);
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(sm);
  const int ldoff8 = (lr + 8 * (lj & 1)) * 4;

  if (blockIdx.x < p.B) prefetch_frame(p.frames + (size_t)blockIdx.x * 12288, sm, tid);
  // ---- prologue: zero every haloed plane once (interiors are overwritten per frame, halos stay zero), weights
  for (int e = tid; e < oX3 - oE0; e += NT) sm[oE0 + e] = 0.f;
  for (int e = tid; e < szW / 2; e += NT) {
    const int ln = e & 31, gg = ln >> 2, tt = ln & 3;
    float x = 0.f, y = 0.f;
    int s = e >> 5;
    if (s < 6) {
      const int ky = s >> 1, kk = s & 1;
      if (tt < 3) {
        x = __ldg(p.w0 + gg * 27 + tt * 9 + ky * 3 + (kk ? 2 : 0));
        if (!kk) y = __ldg(p.w0 + gg * 27 + tt * 9 + ky * 3 + 1);
      }
    } else if ((s -= 6) < 18) {
      const float* w = s < 9 ? p.w1 : p.w2;
      const int tap = s % 9;
      x = __ldg(w + (gg * 8 + tt) * 9 + tap); y = __ldg(w + (gg * 8 + tt + 4) * 9 + tap);
    } else {
      s -= 18;
      const int tap = s >> 1, co = (s & 1) * 8 + gg;
      x = __ldg(p.w3 + (co * 8 + tt) * 9 + tap); y = __ldg(p.w3 + (co * 8 + tt + 4) * 9 + tap);
    }
    sm[oW + 2 * e] = tf32r(x);
    sm[oW + 2 * e + 1] = tf32r(y);
  }
  if (tid < 8) {
    sm[oBias + tid] = __ldg(p.b0 + tid); sm[oBias + 8 + tid] = __ldg(p.b1 + tid); sm[oBias + 16 + tid] = __ldg(p.b2 + tid);
    sm[oBias + 56 + tid] = __ldg(p.bd2 + tid); sm[oBias + 64 + tid] = __ldg(p.bd1 + tid); sm[oBias + 72 + tid] = __ldg(p.bd0 + tid);
  }
  if (tid < 16) { sm[oBias + 24 + tid] = __ldg(p.b3 + tid); sm[oBias + 40 + tid] = __ldg(p.bd3 + tid); }
  for (int e = tid; e < 1024; e += NT) { sm[oHW + hWl1 + e] = __ldg(p.wl1 + e); sm[oHW + hWd4 + e] = __ldg(p.wd4 + e); }
  if (tid < 32) {
    sm[oHW + hBl1 + tid] = __ldg(p.bl1 + tid); sm[oHW + hWl2 + tid] = __ldg(p.wl2 + tid);
    sm[oHW + hB4 + tid] = __ldg(p.b4 + tid); sm[oHW + hBd4 + tid] = __ldg(p.bd4 + tid);
  }
  if (tid == 0) sm[oHW + hBl2] = __ldg(p.bl2);

  for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
    asm volatile("cp.async.wait_all;\n" ::);
    __syncthreads();
    stage_frame(reinterpret_cast<const uint8_t*>(sm + oU8), sm + oX, tid);
    __syncthreads();
    if (n + (int)gridDim.x < p.B) prefetch_frame(p.frames + (size_t)(n + gridDim.x) * 12288, sm, tid);

    // ================= conv0: features.0 (3 -> 8) + ReLU + pool -> e0
    {
      const int x0 = (warp & 3) * 16, r0 = (warp >> 2) * 16;
      float2 w[3][2];
#pragma unroll
      for (int s = 0; s < 6; ++s) w[s >> 1][s & 1] = sWf[(wL0 >> 1) + s * 32 + lane];
      const uint32_t aA = smb + (oX + (r0 * P0 + x0 + lr + 8 * (lj & 1) + (lj >> 1)) * 4) * 4;
      const uint32_t aB = smb + (oX + (r0 * P0 + x0 + lr + 8 * (lj & 1) + 2) * 4) * 4;
      const float bias0 = sm[oBias + 2 * t], bias1 = sm[oBias + 2 * t + 1];
      const int co = 2 * t + odd;
      float* dE = sm + oE0 + (co >> 2) * PL1 + (((r0 >> 1) + 1) * P1 + ((x0 + g) >> 1) + 1) * 4 + (co & 3);
      slide_rows<16, 2>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + i * (P0 * 16));
            ldsm2(a[1][0], a[1][1], aB + i * (P0 * 16));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
            pool_fwd(top, bot, bias0, bias1, odd, [&](int h, float v) { dE[((e >> 1) * P1 + 4 * h) * 4] = tf32r(v); });
          });
    }
    __syncthreads();
    // ================= conv1: features.3 on 32x32 -> e1
    {
      const int x0 = (warp & 1) * 16, r0 = (warp >> 1) * 4;
      float2 w[3][3];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3] = sWf[(wL1 >> 1) + s * 32 + lane];
      const uint32_t aA = smb + (oE0 + (lj >> 1) * PL1 + (r0 * P1 + x0) * 4 + ldoff8) * 4;
      const float bias0 = sm[oBias + 8 + 2 * t], bias1 = sm[oBias + 8 + 2 * t + 1];
      const int co = 2 * t + odd;
      float* dE = sm + oE1 + (co >> 2) * PL2 + (((r0 >> 1) + 1) * P2 + ((x0 + g) >> 1) + 1) * 4 + (co & 3);
      slide_rows<4, 3>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (i * P1 + kx) * 16);
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
            pool_fwd(top, bot, bias0, bias1, odd, [&](int h, float v) { dE[((e >> 1) * P2 + 4 * h) * 4] = tf32r(v); });
          });
    }
    __syncthreads();
    // ================= conv2: features.6 on 16x16 -> e2 (eval mode: Dropout is the identity)
    if (warp < 8) {
      const int r0 = warp * 2;
      float2 w[3][3];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3] = sWf[(wL2 >> 1) + s * 32 + lane];
      const uint32_t aA = smb + (oE1 + (lj >> 1) * PL2 + (r0 * P2) * 4 + ldoff8) * 4;
      const float bias0 = sm[oBias + 16 + 2 * t], bias1 = sm[oBias + 16 + 2 * t + 1];
      slide_rows<2, 3>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (i * P2 + kx) * 16);
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
            pool_fwd(top, bot, bias0, bias1, odd, [&](int h, float v) {
              const int py = warp, px = (g >> 1) + 4 * h, co = 2 * t + odd;
              sm[oE2 + (co >> 2) * PL3 + ((py + 1) * P3 + px + 1) * 4 + (co & 3)] = tf32r(v);
            });
          });
    }
    __syncthreads();
    // ================= conv3: features.10 (8 -> 16) on 8x8 -> e3 (K order for the head + planes 0-3 of the dec[3] operand)
    float4 w4r[4];
    const int rot4 = (tid >> 1) & 3;
    {
      const float4* wr = reinterpret_cast<const float4*>(p.w4 + (tid >> 4) * 256 + (tid & 15) * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) w4r[i] = __ldg(wr + ((i + rot4) & 3));
    }
    if (warp < 8) {
      const int mt = warp >> 1, nt = warp & 1;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t aA = smb + (oE2 + (lj >> 1) * PL3 + ((2 * mt + (lj & 1)) * P3 + lr) * 4) * 4;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        uint32_t a[4];
        ldsm4(a, aA + ((tap / 3) * P3 + tap % 3) * 16);
        const float2 w = sWf[(wL3 >> 1) + (tap * 2 + nt) * 32 + lane];
        mma_tf32(acc, a, __float_as_uint(w.x), __float_as_uint(w.y));
      }
      const int co = nt * 8 + 2 * t + odd;
      const float b0 = sm[oBias + 24 + nt * 8 + 2 * t], b1 = sm[oBias + 24 + nt * 8 + 2 * t + 1];
      const float t0 = acc[0] + b0, t1 = acc[1] + b1, u0 = acc[2] + b0, u1 = acc[3] + b1;
      const float rt = __shfl_xor_sync(0xffffffffu, odd ? t0 : t1, 4), rb = __shfl_xor_sync(0xffffffffu, odd ? u0 : u1, 4);
      const float own = odd ? fmaxf(t1, u1) : fmaxf(t0, u0);
      const float m = fmaxf(fmaxf(own, fmaxf(rt, rb)), 0.f);
      const int py = mt, px = g >> 1;
      sm[oX3 + co * 16 + py * 4 + px] = m;
      sm[oCat3 + (co >> 2) * PL4 + ((py + 1) * P4 + px + 1) * 4 + (co & 3)] = tf32r(m);
    }
    __syncthreads();
    // ================= features.14 (4x4 valid conv = 256 -> 32) + ReLU -> h = embeds[4]
    {
      const int nn = tid >> 4, part = tid & 15;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 aq = w4r[i];
        const float4 bq = *reinterpret_cast<const float4*>(sm + oX3 + part * 16 + ((i + rot4) & 3) * 4);
        s = fmaf(aq.x, bq.x, s); s = fmaf(aq.y, bq.y, s); s = fmaf(aq.z, bq.z, s); s = fmaf(aq.w, bq.w, s);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0) sH[nn] = fmaxf(s + sm[oHW + hB4 + nn], 0.f);
    }
    __syncthreads();
    // ================= crit.1 Linear + ReLU -> v ;  dec[4] (1x1 conv on the bottleneck) -> planes 4-11 of the dec[3] operand
    {
      const int nn = tid >> 4, part = tid & 15;
      const float2 wv = *reinterpret_cast<const float2*>(sm + oHW + hWl1 + nn * 32 + 2 * part);
      const float2 wd = *reinterpret_cast<const float2*>(sm + oHW + hWd4 + nn * 32 + 2 * part);
      const float h0 = sH[2 * part], h1 = sH[2 * part + 1];
      float s = wv.x * h0 + wv.y * h1, d = wd.x * h0 + wd.y * h1;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); d += __shfl_xor_sync(0xffffffffu, d, o); }
      if (part == 0) sV[nn] = fmaxf(s + sm[oHW + hBl1 + nn], 0.f);
      // nearest x4 of a 1x1 map = the value at all 16 pixels (zero outside: the conv's padding); lane `part` writes pixel `part`
      sm[oCat3 + (4 + (nn >> 2)) * PL4 + (((part >> 2) + 1) * P4 + (part & 3) + 1) * 4 + (nn & 3)] = tf32r(d + sm[oHW + hBd4 + nn]);
    }
    __syncthreads();
    // ================= dec[3]: 48 -> 16 on 4x4 : 108 MMAs, K split over 8 warp groups x 2 channel tiles; pred on the side
    {
      const int nt = warp & 1, grp = warp >> 1;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int pix = lr + 8 * (lj & 1), py = pix >> 2, px = pix & 3;
      for (int ks = grp * 7; ks < min(54, grp * 7 + 7); ++ks) {
        const int tap = ks / 6, pg = ks - tap * 6, ky = tap / 3, kx = tap - ky * 3;
        uint32_t a[4];
        ldsm4(a, smb + (oCat3 + (2 * pg + (lj >> 1)) * PL4 + ((py + ky) * P4 + px + kx) * 4) * 4);
        const float2 w = __ldg(pk + (pkD3 >> 1) + (ks * 2 + nt) * 32 + lane);
        mma_tf32(acc, a, __float_as_uint(w.x), __float_as_uint(w.y));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) sm[oScr + ((grp * 2 + nt) * 16 + g + 8 * (q >> 1)) * 8 + 2 * t + (q & 1)] = acc[q];
      if (warp == 15) {          // value head: Linear(32,1) + Sigmoid (eval mode: no dropout)
        const float z = warp_sum(sm[oHW + hWl2 + lane] * sV[lane]) + sm[oHW + hBl2];
        if (lane == 0) p.pred[n] = sigmoidf_(z);
      }
    }
    __syncthreads();
    if (tid < 256) {
      const int nt = tid >> 7, pix = (tid >> 3) & 15, c8 = tid & 7, co = nt * 8 + c8;
      float s = sm[oBias + 40 + co];
#pragma unroll
      for (int grp = 0; grp < 8; ++grp) s += sm[oScr + ((grp * 2 + nt) * 16 + pix) * 8 + c8];
      sm[oO3 + (co >> 2) * PL4 + (((pix >> 2) + 1) * P4 + (pix & 3) + 1) * 4 + (co & 3)] = tf32r(s);
    }
    __syncthreads();
    // ================= dec[2]: cat(e2, up(o3)) 24 -> 8 on 8x8 : 4 row-pair tiles x K split over 4 warp groups
    {
      const int mt = warp & 3, grp = warp >> 2;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int vy0 = 2 * mt + (lj & 1);
      for (int ks = grp * 7; ks < min(27, grp * 7 + 7); ++ks) {
        const int tap = ks / 3, pg = ks - tap * 3, ky = tap / 3, kx = tap - ky * 3;
        uint32_t a[4];
        if (pg == 0) {
          ldsm4(a, smb + (oE2 + (lj >> 1) * PL3 + ((vy0 + ky) * P3 + lr + kx) * 4) * 4);
        } else {
          const int sy = (vy0 + ky + 1) >> 1, sx = (lr + kx + 1) >> 1;
          ldsm4(a, smb + (oO3 + (2 * (pg - 1) + (lj >> 1)) * PL4 + (sy * P4 + sx) * 4) * 4);
        }
        const float2 w = __ldg(pk + (pkD2 >> 1) + ks * 32 + lane);
        mma_tf32(acc, a, __float_as_uint(w.x), __float_as_uint(w.y));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) sm[oScr + 2048 + ((grp * 4 + mt) * 16 + g + 8 * (q >> 1)) * 8 + 2 * t + (q & 1)] = acc[q];
    }
    __syncthreads();
    {
      const int mt = tid >> 7, r = (tid >> 3) & 15, co = tid & 7;
      float s = sm[oBias + 56 + co];
#pragma unroll
      for (int grp = 0; grp < 4; ++grp) s += sm[oScr + 2048 + ((grp * 4 + mt) * 16 + r) * 8 + co];
      const int y = 2 * mt + (r >> 3), x = r & 7;
      sm[oO2 + (co >> 2) * PL3 + ((y + 1) * P3 + x + 1) * 4 + (co & 3)] = tf32r(s);
    }
    __syncthreads();
    // ================= dec[1]: cat(e1, up(o2)) 16 -> 8 on 16x16 : 8 warps x 2 rows
    if (warp < 8) {
      const int r0 = warp * 2;
      float2 w[3][6];
#pragma unroll
      for (int s = 0; s < 18; ++s) w[s / 6][s % 6] = __ldg(pk + (pkD1 >> 1) + s * 32 + lane);
      const uint32_t aE = smb + (oE1 + (lj >> 1) * PL2 + (r0 * P2) * 4 + ldoff8) * 4;
      const int vx = lr + 8 * (lj & 1);
      const float bias0 = sm[oBias + 64 + 2 * t], bias1 = sm[oBias + 64 + 2 * t + 1];
      slide_rows<2, 6>(
          w,
          [&](int i, uint32_t(&a)[6][4]) {
            const int sy = (r0 + i + 1) >> 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              ldsm4(a[2 * kx], aE + (i * P2 + kx) * 16);
              ldsm4(a[2 * kx + 1], smb + (oO2 + (lj >> 1) * PL3 + (sy * P3 + ((vx + kx + 1) >> 1)) * 4) * 4);
            }
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int y = r0 + e + r, x = g + 8 * (q >> 1), ch = 2 * t + (q & 1);
                sm[oO1 + (ch >> 2) * PL2 + ((y + 1) * P2 + x + 1) * 4 + (ch & 3)] = tf32r((r ? bot[q] : top[q]) + ((q & 1) ? bias1 : bias0));
              }
          });
    }
    __syncthreads();
    // ================= dec[0]: cat(e0, up(o1)) 16 -> 8 on 32x32 -> o0 in global memory (NHWC)
    {
      const int x0 = (warp & 1) * 16, r0 = (warp >> 1) * 4;
      float2 w[3][6];
#pragma unroll
      for (int s = 0; s < 18; ++s) w[s / 6][s % 6] = __ldg(pk + (pkD0 >> 1) + s * 32 + lane);
      const uint32_t aE = smb + (oE0 + (lj >> 1) * PL1 + (r0 * P1 + x0) * 4 + ldoff8) * 4;
      const int vx = x0 + lr + 8 * (lj & 1);
      const float bias0 = sm[oBias + 72 + 2 * t], bias1 = sm[oBias + 72 + 2 * t + 1];
      float* dO = p.o0 + ((size_t)n * 1024 + r0 * 32 + x0 + g) * 8 + 2 * t;
      slide_rows<4, 6>(
          w,
          [&](int i, uint32_t(&a)[6][4]) {
            const int sy = (r0 + i + 1) >> 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              ldsm4(a[2 * kx], aE + (i * P1 + kx) * 16);
              ldsm4(a[2 * kx + 1], smb + (oO1 + (lj >> 1) * PL2 + (sy * P2 + ((vx + kx + 1) >> 1)) * 4) * 4);
            }
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float v0 = (r ? bot[2 * h] : top[2 * h]) + bias0, v1 = (r ? bot[2 * h + 1] : top[2 * h + 1]) + bias1;
                *reinterpret_cast<float2*>(dO + ((e + r) * 32 + 8 * h) * 8) = make_float2(tf32r(v0), tf32r(v1));
              }
          });
    }
    // the next iteration's barrier after cp.async.wait_all separates these reads from the next frame's writes
  }
  asm volatile("cp.async.wait_all;\n" ::);
}

// =====================================================================================================================
// masker[0..3] (reference nets.py:488-491, 519-523) in one kernel: cat(X, ups(o0)) -> Conv2d(11,16,3,1,1) -> LeakyReLU(0.01)
// -> Conv2d(16,1,3,1,1) -> Sigmoid [-> >= threshold].  62 % of the inference MACs; unfused, the 16-channel 64x64 map (256 KB
// per frame) makes a round trip through HBM.  Here the map exists only as an 18-row band in shared memory: four bands of 16
// mask rows per frame, the first conv computes the band's 18 rows (one row of overlap each side), the second consumes them.
namespace mk {
using namespace cf;
constexpr int NT = 512;
constexpr int P0 = 66, SX = 66 * 66 * 4;
constexpr int P1 = 34, PL1 = 34 * 34 * 4;
constexpr int PLB = 18 * 66 * 4;                      // one 4-channel plane of the band: 18 rows x 66 pixels
constexpr int oX = 0, oO0 = oX + SX, oBand = oO0 + 2 * PL1, oU8 = oBand + 4 * PLB;
constexpr int oW0 = oU8 + 3072;                      // masker.0 fragments: step (ky*5 + kk)*2 + nt
constexpr int oW2 = oW0 + 30 * 64;                   // masker.2 fragments: step ky*6 + kx*2 + pg
constexpr int oBias = oW2 + 18 * 64;                 // bm0[16] bm2[1]
constexpr int SMEM_FLOATS = oBias + 32;
static_assert(SMEM_FLOATS * 4 <= 227 * 1024, "shared memory budget");

struct Params {
  const uint8_t* frames;
  const float* o0;
  const float *w0, *b0, *w2, *b2;
  float* mask;
  uint8_t* hard;
  float thresh;
  int B;
};

__device__ __forceinline__ void prefetch(const Params& p, int n, float* sm, int tid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(sm);
  const uint8_t* src = p.frames + (size_t)n * 12288;
  for (int c = tid; c < 768; c += NT)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + (oU8 * 4) + c * 16), "l"(src + c * 16));
  asm volatile("cp.async.commit_group;\n" ::);
}
// o0 [32][32][8] fp32 (already TF32-rounded by its producer) -> two haloed half-planes, 16 bytes at a time
__device__ __forceinline__ void fetch_o0(const Params& p, int n, float* sm, int tid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(sm);
  const float* src = p.o0 + (size_t)n * 8192;
  for (int c = tid; c < 2048; c += NT) {
    const int pix = c >> 1, half = c & 1, y = pix >> 5, x = pix & 31;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + (oO0 + half * PL1 + ((y + 1) * P1 + x + 1) * 4) * 4),
                 "l"(src + c * 4));
  }
  asm volatile("cp.async.commit_group;\n" ::);
}

// masker.0 on R rows of the band starting at band row rb0 (absolute mask row = 16*band - 1 + band row)
template <int R>
__device__ __forceinline__ void m0_rows(float* sm, uint32_t smb, const float2 (&w)[3][5], int band, int rb0, int x0, int nt, int lane,
                                        float bias0, float bias1) {
  const int lj = lane >> 3, lr = lane & 7, g = lane >> 2, t = lane & 3;
  const int vr0 = 16 * band - 1 + rb0;               // absolute output row of band row rb0; haloed virtual input row = vr0 + i
  const uint32_t aA = smb + (oX + (x0 + lr + 8 * (lj & 1) + (lj >> 1)) * 4) * 4;
  const uint32_t aB = smb + (oX + (x0 + lr + 8 * (lj & 1) + 2) * 4) * 4;
  const int vx = x0 + lr + 8 * (lj & 1);
  slide_rows<R, 5>(
      w,
      [&](int i, uint32_t(&a)[5][4]) {
        int hv = vr0 + i;                             // haloed virtual row 0..65; rows of the out-of-frame band rows are clamped
        hv = hv < 0 ? 0 : (hv > 65 ? 65 : hv);        // (their results are discarded below)
        ldsm4(a[0], aA + hv * (P0 * 16));
        ldsm2(a[1][0], a[1][1], aB + hv * (P0 * 16));
        a[1][2] = a[1][3] = 0u;
        const int sy = (hv + 1) >> 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          ldsm4(a[2 + kx], smb + (oO0 + (lj >> 1) * PL1 + (sy * P1 + ((vx + kx + 1) >> 1)) * 4) * 4);
      },
      [&](int e, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int rb = rb0 + e + r, ya = 16 * band - 1 + rb;
          const bool inside = ya >= 0 && ya < 64;     // rows -1 and 64 are the second conv's zero padding
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int x = x0 + g + 8 * (q >> 1), co = nt * 8 + 2 * t + (q & 1);
            float v = (r ? bot[q] : top[q]) + ((q & 1) ? bias1 : bias0);
            v = v > 0.f ? v : v * kLeakySlope;
            sm[oBand + (co >> 2) * PLB + (rb * P0 + x + 1) * 4 + (co & 3)] = inside ? tf32r(v) : 0.f;
          }
        }
      });
}

__global__ void __launch_bounds__(NT, 1) masker_fused_kernel(const Params p) {
  extern __shared__ __align__(128) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int lj = lane >> 3, lr = lane & 7;
  const float2* sW0 = reinterpret_cast<const float2*>(sm + oW0);
  const float2* sW2 = reinterpret_cast<const float2*>(sm + oW2);
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(sm);

  if (blockIdx.x < p.B) { prefetch(p, blockIdx.x, sm, tid); fetch_o0(p, blockIdx.x, sm, tid); }
  // ---- prologue: zero the operand planes (halos stay zero), weight fragments
  for (int e = tid; e < 4 * PLB; e += NT) sm[oBand + e] = 0.f;
  for (int e = tid; e < 2 * PL1 / 4; e += NT) {      // only the halo ring must be zero; cp.async is filling the interior
    const int h = e / (PL1 / 4), q = e - h * (PL1 / 4), y = q / P1, x = q - y * P1;
    if (y == 0 || y == 33 || x == 0 || x == 33) *reinterpret_cast<float4*>(sm + oO0 + h * PL1 + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int e = tid; e < 30 * 32; e += NT) {
    const int ln = e & 31, gg = ln >> 2, tt = ln & 3, s = e >> 5, nt = s & 1, kk = (s >> 1) % 5, ky = (s >> 1) / 5;
    const float* w = p.w0 + (size_t)(nt * 8 + gg) * 99;          // [16][11][3][3]
    float x = 0.f, y = 0.f;
    if (kk == 0) { if (tt < 3) { x = __ldg(w + tt * 9 + ky * 3); y = __ldg(w + tt * 9 + ky * 3 + 1); } }
    else if (kk == 1) { if (tt < 3) x = __ldg(w + tt * 9 + ky * 3 + 2); }
    else { x = __ldg(w + (3 + tt) * 9 + ky * 3 + kk - 2); y = __ldg(w + (7 + tt) * 9 + ky * 3 + kk - 2); }
    sm[oW0 + 2 * e] = tf32r(x); sm[oW0 + 2 * e + 1] = tf32r(y);
  }
  for (int e = tid; e < 18 * 32; e += NT) {
    const int ln = e & 31, gg = ln >> 2, tt = ln & 3, s = e >> 5, pg = s & 1, tap = (s / 6) * 3 + (s % 6) / 2;
    float x = 0.f, y = 0.f;
    if (gg == 0) { x = __ldg(p.w2 + (pg * 8 + tt) * 9 + tap); y = __ldg(p.w2 + (pg * 8 + tt + 4) * 9 + tap); }
    sm[oW2 + 2 * e] = tf32r(x); sm[oW2 + 2 * e + 1] = tf32r(y);
  }
  if (tid < 16) sm[oBias + tid] = __ldg(p.b0 + tid);
  if (tid == 16) sm[oBias + 16] = __ldg(p.b2);

  for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
    asm volatile("cp.async.wait_all;\n" ::);
    __syncthreads();
    inf::stage_frame(reinterpret_cast<const uint8_t*>(sm + oU8), sm + oX, tid);
    __syncthreads();
    if (n + (int)gridDim.x < p.B) prefetch(p, n + gridDim.x, sm, tid);
    for (int band = 0; band < 4; ++band) {
      // ---- masker.0 + LeakyReLU on the band's 18 rows: 4 strips x 2 channel tiles x 2 row segments (10 + 8 rows)
      {
        const int x0 = (warp & 3) * 16, nt = (warp >> 2) & 1, seg = warp >> 3;
        float2 w[3][5];
#pragma unroll
        for (int s = 0; s < 15; ++s) w[s / 5][s % 5] = sW0[(s * 2 + nt) * 32 + lane];
        const float bias0 = sm[oBias + nt * 8 + 2 * t], bias1 = sm[oBias + nt * 8 + 2 * t + 1];
        if (seg == 0) m0_rows<10>(sm, smb, w, band, 0, x0, nt, lane, bias0, bias1);
        else m0_rows<8>(sm, smb, w, band, 10, x0, nt, lane, bias0, bias1);
      }
      __syncthreads();
      // ---- masker.2 + Sigmoid (+ threshold) on the band's 16 mask rows: 4 strips x 4 segments of 4 rows
      {
        const int x0 = (warp & 3) * 16, r0 = (warp >> 2) * 4;
        float2 w[3][6];
#pragma unroll
        for (int s = 0; s < 18; ++s) w[s / 6][s % 6] = sW2[s * 32 + lane];
        const uint32_t aA = smb + (oBand + (lj >> 1) * PLB + (r0 * P0 + x0 + lr + 8 * (lj & 1)) * 4) * 4;
        const float b2 = sm[oBias + 16];
        float* dM = p.mask + (size_t)n * 4096 + (16 * band + r0) * 64 + x0 + g;
        uint8_t* dH = p.hard ? p.hard + (size_t)n * 4096 + (16 * band + r0) * 64 + x0 + g : nullptr;
        slide_rows<4, 6>(
            w,
            [&](int i, uint32_t(&a)[6][4]) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                ldsm4(a[2 * kx], aA + ((i * P0 + kx) * 4) * 4);
                ldsm4(a[2 * kx + 1], aA + (2 * PLB + (i * P0 + kx) * 4) * 4);
              }
            },
            [&](int e, const float(&top)[4], const float(&bot)[4]) {
              if (t == 0) {                             // column 0 of the 8-wide tile is the one real output channel
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const float z = sigmoidf_((r ? bot[2 * h] : top[2 * h]) + b2);
                    dM[(e + r) * 64 + 8 * h] = z;
                    if (dH) dH[(e + r) * 64 + 8 * h] = z >= p.thresh;
                  }
              }
            });
      }
      __syncthreads();
    }
    if (n + (int)gridDim.x < p.B) fetch_o0(p, n + gridDim.x, sm, tid);   // all reads of this frame's o0 planes are done
  }
  asm volatile("cp.async.wait_all;\n" ::);
}
}  // namespace mk

}  // namespace inf
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_masker_fused(const uint8_t* frames, const float* o0, int32_t B, const float* wm0, const float* bm0,
                                const float* wm2, const float* bm2, float thresh, float* mask, uint8_t* hard, void* stream) {
  CGS_REQUIRE(frames && o0 && wm0 && bm0 && wm2 && bm2 && mask && B > 0, "masker_fused: bad args");
  CGS_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)o0 & 15) == 0, "masker_fused: frames and o0 must be 16-byte aligned");
  inf::mk::Params p;
  p.frames = frames; p.o0 = o0; p.w0 = wm0; p.b0 = bm0; p.w2 = wm2; p.b2 = bm2; p.mask = mask; p.hard = hard; p.thresh = thresh;
  p.B = B;
  const int sms = device_sms();
  cudaFuncSetAttribute(inf::mk::masker_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, inf::mk::SMEM_FLOATS * 4);
  const int per = (B + sms - 1) / sms, grid = (B + per - 1) / per;
  inf::mk::masker_fused_kernel<<<grid, inf::mk::NT, inf::mk::SMEM_FLOATS * 4, (cudaStream_t)stream>>>(p);
  return check_launch("masker_fused");
}

extern "C" int cgs_infer_pack_floats(void) { return inf::PACK_FLOATS; }

extern "C" int cgs_infer_pack_decoder(const float* wd3, const float* wd2, const float* wd1, const float* wd0, float* pack,
                                      void* stream) {
  CGS_REQUIRE(wd3 && wd2 && wd1 && wd0 && pack, "infer_pack_decoder: bad args");
  inf::pack_decoder_kernel<<<(inf::PACK_FLOATS / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wd3, wd2, wd1, wd0, pack);
  return check_launch("infer_pack_decoder");
}

extern "C" int cgs_infer_fused(const uint8_t* frames, int32_t B, const cgs_critic_weights* cw, const float* wd4, const float* bd4,
                               const float* bd3, const float* bd2, const float* bd1, const float* bd0, const float* pack,
                               float* pred, float* o0, void* stream) {
  CGS_REQUIRE(frames && cw && wd4 && bd4 && bd3 && bd2 && bd1 && bd0 && pack && pred && o0 && B > 0, "infer_fused: bad args");
  CGS_REQUIRE(((uintptr_t)frames & 15) == 0 && ((uintptr_t)pack & 7) == 0 && ((uintptr_t)o0 & 7) == 0, "infer_fused: alignment");
  inf::Params p;
  p.frames = frames; p.B = B;
  p.w0 = cw->w0; p.b0 = cw->b0; p.w1 = cw->w1; p.b1 = cw->b1; p.w2 = cw->w2; p.b2 = cw->b2; p.w3 = cw->w3; p.b3 = cw->b3;
  p.w4 = cw->w4; p.b4 = cw->b4; p.wl1 = cw->wl1; p.bl1 = cw->bl1; p.wl2 = cw->wl2; p.bl2 = cw->bl2;
  p.wd4 = wd4; p.bd4 = bd4; p.bd3 = bd3; p.bd2 = bd2; p.bd1 = bd1; p.bd0 = bd0; p.pack = pack; p.pred = pred; p.o0 = o0;
  const int sms = device_sms();
  cudaFuncSetAttribute(inf::infer_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, inf::SMEM_FLOATS * 4);
  const int per = (B + sms - 1) / sms, grid = (B + per - 1) / per;
  inf::infer_fused_kernel<<<grid, inf::NT, inf::SMEM_FLOATS * 4, (cudaStream_t)stream>>>(p);
  return check_launch("infer_fused");
}
