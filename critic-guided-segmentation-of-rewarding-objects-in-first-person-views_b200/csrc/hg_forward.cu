// The Hourglass forward as ONE persistent kernel, bf16 tensor-core operands with fp32 accumulation, chfak = 1:
//   uint8 frame -> /255 (+ shift_batch roll) -> NewCritic.forward(collect=True) (reference nets.py:197-212; train mode:
//   dropout masks applied) -> UnetDecoder.forward (nets.py:494-523): dec[4] .. dec[0], masker.0 + LeakyReLU, masker.2 +
//   Sigmoid [-> >= threshold, main.py:1164].
// It serves both `-process` inference (reference main.py:1139-1164; eval mode, no tape) and the forward half of a
// segmentation_training step (main.py:364, 391; train mode): there it also leaves the per-frame TAPE (every skip / decoder
// activation as haloed bf16 planes, 54 KB per frame) that cgs_hg_backward consumes, so nothing is recomputed.
//
// One CTA (16 warps) owns a frame at a time; every activation lives in shared memory (hg_common.cuh has the layout rules);
// HBM traffic per frame = 12 KB frame in, 16 KB mask (+4 KB hard mask, +54 KB tape in training) out.  The 16-channel 64x64
// masker map (256 KB per frame in fp32) only ever exists as an 18-row bf16 band.
#include <string.h>
#include "hg_common.cuh"

namespace cgs {
namespace hg {

// ---- shared memory map of the forward kernel (byte offsets)
constexpr int fX = TAPE;                            // pair-duplicated frame
constexpr int fBand = fX + PBX;                     // masker.0 band: 2 planes x 10 rows; before the bands: split-K scratch
constexpr int PXP = 68;                             // pitch of the partial-product rows: column x + 1, columns 0 and 65.. zero
constexpr int fP = fBand + 2 * PLB8;                // masker.2 partial products, fp32 [10 rows][9 taps][PXP]
constexpr int fE2D = fP + 10 * 9 * PXP * 4;         // e2 * dropout mask (operand of features.10)
constexpr int fU8 = fE2D + PB3;                     // raw frame bytes (prefetched)
constexpr int fX3 = fU8 + 12288;                    // e3 * mask in the 4x4 conv's K order, fp32 [256]
constexpr int fV = fX3 + 1024;                      // v[32] fp32
constexpr int fM2 = fV + 128, fM3 = fM2 + 2048, fMV = fM3 + 1024;   // dropout masks fp32 [512] [256] [32]
constexpr int fW = fMV + 128;                       // weight fragments, steps [0, F_SMEM_STEPS) | [F_PT, F_PT + 2)
constexpr int fBias = fW + (F_SMEM_STEPS + 2) * 256;      // fp32: b0[8] b1[8] b2[8] b3[16] bd3[16] bd2[8] bd1[8] bd0[8] bm0[16] bm2[1]
constexpr int bB0 = 0, bB1 = 8, bB2 = 16, bB3 = 24, bD3 = 40, bD2 = 56, bD1 = 64, bD0 = 72, bM0 = 80, bM2 = 96;
constexpr int fHW = fBias + 512;                    // fp32: wl1[1024] bl1[32] wl2[32] bl2[4] b4[32] wd4[1024] bd4[32]
constexpr int hWl1 = 0, hBl1 = 1024, hWl2 = 1056, hBl2 = 1088, hB4 = 1092, hWd4 = 1124, hBd4 = 2148, szHW = 2180;
constexpr int F_SMEM = fHW + szHW * 4;
static_assert(F_SMEM <= 227 * 1024, "forward kernel: shared memory budget");
static_assert(2 * PLB8 >= 4096 * 4, "the split-K scratch aliases the band");
static_assert(fX % 16 == 0 && fBand % 16 == 0 && fE2D % 16 == 0 && fU8 % 16 == 0 && fW % 16 == 0, "16-byte alignment");

struct FwdParams {
  const uint8_t* frames;
  const uint2* pack;
  const float *b0, *b1, *b2, *b3, *w4, *b4, *wl1, *bl1, *wl2, *bl2;
  const float *wd4, *bd4, *bd3, *bd2, *bd1, *bd0, *bm0, *bm2;
  const float *m2, *m3, *mv;          // forced dropout masks (train), or NULL
  unsigned long long seed;
  unsigned long long* rng_state;     // != NULL: masks drawn in the kernel, stream of cgs_dropout_masks
  float p_drop, keep;
  const int* roll_dev;
  int B, roll, train;
  float thresh;
  float* pred;
  float* z;
  uint8_t* hard;
  uint8_t* tape;
};

__global__ void hg_pack_kernel(const PackSrc p, uint2* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NSTEPS * 32) return;
  const int s = e >> 5, lane = e & 31, g = lane >> 2, t = lane & 3;
  uint2 v;
  v.x = pack_bf16(pack_wk(p, s, 2 * t, g), pack_wk(p, s, 2 * t + 1, g));
  v.y = pack_bf16(pack_wk(p, s, 2 * t + 8, g), pack_wk(p, s, 2 * t + 9, g));
  out[e] = v;
}

__device__ long long* g_hgf_trace = nullptr;

__global__ void __launch_bounds__(NT, 1) hg_forward_kernel(const FwdParams p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, odd = g & 1;
  const int lj = lane >> 3, lr = lane & 7, pixoff = lr + 8 * (lj & 1), chunk = lj >> 1;
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(smraw);
  const uint2* sWf = reinterpret_cast<const uint2*>(smraw + fW);
  float* sBias = reinterpret_cast<float*>(smraw + fBias);
  float* sHW = reinterpret_cast<float*>(smraw + fHW);
  float* sH = reinterpret_cast<float*>(smraw + tH);
  float* sV = reinterpret_cast<float*>(smraw + fV);
  float* sX3 = reinterpret_cast<float*>(smraw + fX3);
  float* sM2 = reinterpret_cast<float*>(smraw + fM2);
  float* sM3 = reinterpret_cast<float*>(smraw + fM3);
  float* sMV = reinterpret_cast<float*>(smraw + fMV);
  float* sScr = reinterpret_cast<float*>(smraw + fBand);
  __nv_bfloat16* sE0 = reinterpret_cast<__nv_bfloat16*>(smraw + tE0);
  __nv_bfloat16* sE1 = reinterpret_cast<__nv_bfloat16*>(smraw + tE1);
  __nv_bfloat16* sE2 = reinterpret_cast<__nv_bfloat16*>(smraw + tE2);
  __nv_bfloat16* sE2D = reinterpret_cast<__nv_bfloat16*>(smraw + fE2D);
  __nv_bfloat16* sC3 = reinterpret_cast<__nv_bfloat16*>(smraw + tC3);
  __nv_bfloat16* sO3 = reinterpret_cast<__nv_bfloat16*>(smraw + tO3);
  __nv_bfloat16* sO2 = reinterpret_cast<__nv_bfloat16*>(smraw + tO2);

  long long* trace = blockIdx.x == 0 ? g_hgf_trace : nullptr;
  int fr = 0;
  const unsigned long long rng_call = p.rng_state ? p.rng_state[0] : 0ull;
  if (blockIdx.x < p.B) {
    const uint8_t* src = p.frames + (size_t)blockIdx.x * 12288;
    for (int c = tid; c < 768; c += NT) cp_async16(smb + fU8 + c * 16, src + c * 16);
    cp_async_commit();
  }
  // ---- prologue: zero every plane once (interiors are overwritten per frame, halos stay zero), weights, biases
  {
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int e = tid; e < fU8 / 16; e += NT) reinterpret_cast<uint4*>(smraw)[e] = z4;
    const uint4* src = reinterpret_cast<const uint4*>(p.pack);
    for (int e = tid; e < F_SMEM_STEPS * 16; e += NT) reinterpret_cast<uint4*>(smraw + fW)[e] = __ldg(src + e);
    if (tid < 32) reinterpret_cast<uint4*>(smraw + fW + F_SMEM_STEPS * 256)[tid] = __ldg(src + F_PT * 16 + tid);
  }
  if (tid < 8) {
    sBias[bB0 + tid] = __ldg(p.b0 + tid); sBias[bB1 + tid] = __ldg(p.b1 + tid); sBias[bB2 + tid] = __ldg(p.b2 + tid);
    sBias[bD2 + tid] = __ldg(p.bd2 + tid); sBias[bD1 + tid] = __ldg(p.bd1 + tid); sBias[bD0 + tid] = __ldg(p.bd0 + tid);
  }
  if (tid < 16) { sBias[bB3 + tid] = __ldg(p.b3 + tid); sBias[bD3 + tid] = __ldg(p.bd3 + tid); sBias[bM0 + tid] = __ldg(p.bm0 + tid); }
  if (tid == 0) { sBias[bM2] = __ldg(p.bm2); sHW[hBl2] = __ldg(p.bl2); }
  for (int e = tid; e < 1024; e += NT) { sHW[hWl1 + e] = __ldg(p.wl1 + e); sHW[hWd4 + e] = __ldg(p.wd4 + e); }
  if (tid < 32) {
    sHW[hBl1 + tid] = __ldg(p.bl1 + tid); sHW[hWl2 + tid] = __ldg(p.wl2 + tid);
    sHW[hB4 + tid] = __ldg(p.b4 + tid); sHW[hBd4 + tid] = __ldg(p.bd4 + tid);
  }
  __syncthreads();                                     // the zero fill above must not race the mask stores below
  if (!p.train || (!p.m2 && !p.rng_state)) {         // eval mode / no dropout: identity masks, written once
    sM2[tid] = 1.f;
    if (tid < 256) sM3[tid] = 1.f;
    if (tid < 32) sMV[tid] = 1.f;
  }
  int roll = p.roll_dev ? *p.roll_dev : p.roll;
  roll = ((roll % 64) + 64) & 63;

  for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
    cp_async_wait_all();
    if (p.tape && tid == 0) bulk_wait_read();          // the previous frame's tape has left shared memory
    __syncthreads();
    HG_MARK(0);
    // ================= frame bytes have landed -> pair-duplicated bf16 plane (rows 1..64); this frame's dropout masks
    stage_rows(smraw + fU8, smraw + fX + PX * 16, 0, 64, roll, tid);
    if (p.train) {
      if (p.rng_state) draw_masks3(p.seed, rng_call, p.B, n, p.p_drop, p.keep, sM2, sM3, sMV, tid);
      else if (p.m2) {
        if (tid < 128) reinterpret_cast<float4*>(sM2)[tid] = __ldg(reinterpret_cast<const float4*>(p.m2 + (size_t)n * 512) + tid);
        else if (tid < 192) reinterpret_cast<float4*>(sM3)[tid - 128] = __ldg(reinterpret_cast<const float4*>(p.m3 + (size_t)n * 256) + tid - 128);
        else if (tid < 200) reinterpret_cast<float4*>(sMV)[tid - 192] = __ldg(reinterpret_cast<const float4*>(p.mv + (size_t)n * 32) + tid - 192);
      }
    }
    __syncthreads();
    if (n + (int)gridDim.x < p.B) {
      const uint8_t* src = p.frames + (size_t)(n + gridDim.x) * 12288;
      for (int c = tid; c < 768; c += NT) cp_async16(smb + fU8 + c * 16, src + c * 16);
      cp_async_commit();
    }

    HG_MARK(1);
    // ================= features.0 (3 -> 8) + ReLU + pool -> e0 : 4 strips x 4 segments of 16 rows, one MMA per filter row
    {
      const int x0 = (warp & 3) * 16, r0 = (warp >> 2) * 16;
      uint2 w[3][1][1];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) w[ky][0][0] = sWf[(F_C0 + ky) * 32 + lane];
      const uint32_t aA = smb + fX + (uint32_t)((r0 * PX + x0 + pixoff + 2 * chunk) * 16);
      const float bias0 = sBias[bB0 + 2 * t], bias1 = sBias[bB0 + 2 * t + 1];
      const int co = 2 * t + odd;
      __nv_bfloat16* dE = sE0 + (((r0 >> 1) + 1) * P1 + ((x0 + g) >> 1) + 1) * 8 + co;
      slide_bf<16, 1, 1>(
          w, [&](int i, uint32_t(&a)[1][4]) { ldsm4(a[0], aA + (uint32_t)(i * (PX * 16))); },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
            pool_fwd(top, bot, bias0, bias1, odd, [&](int h, float v) { dE[((e >> 1) * P1 + 4 * h) * 8] = __float2bfloat16_rn(v); });
          });
    }
    __syncthreads();
    HG_MARK(2);
    // ================= features.3 (8 -> 8) on 32x32 -> e1 : 2 strips x 8 segments of 4 rows
    {
      const int x0 = (warp & 1) * 16, r0 = (warp >> 1) * 4;
      uint2 w[3][2][1];
#pragma unroll
      for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(F_C1 + s) * 32 + lane];
      const uint32_t aA = smb + tE0 + (uint32_t)((r0 * P1 + x0 + pixoff + chunk) * 16);
      const uint32_t aB = smb + tE0 + (uint32_t)((r0 * P1 + x0 + pixoff + 2) * 16);
      const float bias0 = sBias[bB1 + 2 * t], bias1 = sBias[bB1 + 2 * t + 1];
      const int co = 2 * t + odd;
      __nv_bfloat16* dE = sE1 + (((r0 >> 1) + 1) * P2 + ((x0 + g) >> 1) + 1) * 8 + co;
      slide_bf<4, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P1 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P1 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
            pool_fwd(top, bot, bias0, bias1, odd, [&](int h, float v) { dE[((e >> 1) * P2 + 4 * h) * 8] = __float2bfloat16_rn(v); });
          });
    }
    __syncthreads();
    HG_MARK(3);
    // ================= features.6 (8 -> 8) on 16x16 -> e2 (skip, pre-dropout) and e2 * mask (operand of features.10)
    if (warp < 8) {
      const int r0 = warp * 2;
      uint2 w[3][2][1];
#pragma unroll
      for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(F_C2 + s) * 32 + lane];
      const uint32_t aA = smb + tE1 + (uint32_t)((r0 * P2 + pixoff + chunk) * 16);
      const uint32_t aB = smb + tE1 + (uint32_t)((r0 * P2 + pixoff + 2) * 16);
      const float bias0 = sBias[bB2 + 2 * t], bias1 = sBias[bB2 + 2 * t + 1];
      slide_bf<2, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P2 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P2 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int, int, const float(&top)[4], const float(&bot)[4]) {
            pool_fwd(top, bot, bias0, bias1, odd, [&](int h, float v) {
              const int py = warp, px = (g >> 1) + 4 * h, co = 2 * t + odd, o = ((py + 1) * P3 + px + 1) * 8 + co;
              sE2[o] = __float2bfloat16_rn(v);
              sE2D[o] = __float2bfloat16_rn(v * sM2[(py * 8 + px) * 8 + co]);
            });
          });
    }
    __syncthreads();
    HG_MARK(4);
    // ================= features.10 (8 -> 16) on 8x8 -> e3: skip planes (pre-dropout) + head operand (K order, * mask)
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
#pragma unroll
      for (int tp = 0; tp < 5; ++tp) {
        const int tap = min(2 * tp + chunk, 8), ky = tap / 3, kx = tap - 3 * ky;
        uint32_t a[4];
        ldsm4(a, smb + fE2D + (uint32_t)(((2 * mt + (lj & 1) + ky) * P3 + lr + kx) * 16));
        const uint2 w = sWf[(F_C3 + tp * 2 + nt) * 32 + lane];
        mma_bf16(acc, a, w.x, w.y);
      }
      const int co = nt * 8 + 2 * t + odd;
      const float b0 = sBias[bB3 + nt * 8 + 2 * t], b1 = sBias[bB3 + nt * 8 + 2 * t + 1];
      const float t0 = acc[0] + b0, t1 = acc[1] + b1, u0 = acc[2] + b0, u1 = acc[3] + b1;
      const float rt = __shfl_xor_sync(0xffffffffu, odd ? t0 : t1, 4), rb = __shfl_xor_sync(0xffffffffu, odd ? u0 : u1, 4);
      const float own = odd ? fmaxf(t1, u1) : fmaxf(t0, u0);
      const float m = fmaxf(fmaxf(own, fmaxf(rt, rb)), 0.f);
      const int py = mt, px = g >> 1, pp = py * 4 + px;
      sX3[co * 16 + pp] = m * sM3[pp * 16 + co];
      sC3[(co >> 3) * (PB4 / 2) + ((py + 1) * P4 + px + 1) * 8 + (co & 7)] = __float2bfloat16_rn(m);
    }
    __syncthreads();
    HG_MARK(5);
    // ================= features.14 (4x4 valid conv = 256 -> 32) + ReLU -> h = embeds[4]
    {
      const int nn = tid >> 4, part = tid & 15;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 aq = w4r[i];
        const float4 bq = *reinterpret_cast<const float4*>(sX3 + part * 16 + ((i + rot4) & 3) * 4);
        s = fmaf(aq.x, bq.x, s); s = fmaf(aq.y, bq.y, s); s = fmaf(aq.z, bq.z, s); s = fmaf(aq.w, bq.w, s);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0) sH[nn] = fmaxf(s + sHW[hB4 + nn], 0.f);
    }
    __syncthreads();
    // ================= crit.1 Linear + ReLU -> v ; dec[4] (1x1 conv on the bottleneck), nearest x4 -> planes 2-5 of the dec[3] operand
    {
      const int nn = tid >> 4, part = tid & 15;
      const float2 wv = *reinterpret_cast<const float2*>(sHW + hWl1 + nn * 32 + 2 * part);
      const float2 wd = *reinterpret_cast<const float2*>(sHW + hWd4 + nn * 32 + 2 * part);
      const float h0 = sH[2 * part], h1 = sH[2 * part + 1];
      float s = wv.x * h0 + wv.y * h1, d = wd.x * h0 + wd.y * h1;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); d += __shfl_xor_sync(0xffffffffu, d, o); }
      if (part == 0) sV[nn] = fmaxf(s + sHW[hBl1 + nn], 0.f);
      sC3[(2 + (nn >> 3)) * (PB4 / 2) + (((part >> 2) + 1) * P4 + (part & 3) + 1) * 8 + (nn & 7)] = __float2bfloat16_rn(d + sHW[hBd4 + nn]);
    }
    __syncthreads();
    HG_MARK(6);
    // ================= dec[3]: 48 -> 16 on 4x4: 27 k-steps split over 8 warp groups x 2 channel tiles; pred on the side
    {
      const int nt = warp & 1, grp = warp >> 1;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int py = pixoff >> 2, px = pixoff & 3;
      for (int ks = grp * 4; ks < min(27, grp * 4 + 4); ++ks) {
        const int tap = ks / 3, kc = ks - tap * 3, ky = tap / 3, kx = tap - ky * 3;
        uint32_t a[4];
        ldsm4(a, smb + tC3 + (uint32_t)((2 * kc + chunk) * PB4 + ((py + ky) * P4 + px + kx) * 16));
        const uint2 w = __ldg(p.pack + (F_D3 + (tap * 3 + kc) * 2 + nt) * 32 + lane);
        mma_bf16(acc, a, w.x, w.y);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) sScr[((grp * 2 + nt) * 16 + g + 8 * (q >> 1)) * 8 + 2 * t + (q & 1)] = acc[q];
      if (warp == 15) {          // value head: Dropout, Linear(32,1), Sigmoid
        const float zz = warp_sum(sHW[hWl2 + lane] * sV[lane] * sMV[lane]) + sHW[hBl2];
        if (lane == 0) p.pred[n] = sigmoidf_(zz);
      }
    }
    __syncthreads();
    if (tid < 256) {
      const int nt = tid >> 7, pix = (tid >> 3) & 15, c8 = tid & 7, co = nt * 8 + c8;
      float s = sBias[bD3 + co];
#pragma unroll
      for (int grp = 0; grp < 8; ++grp) s += sScr[((grp * 2 + nt) * 16 + pix) * 8 + c8];
      sO3[nt * (PB4 / 2) + (((pix >> 2) + 1) * P4 + (pix & 3) + 1) * 8 + c8] = __float2bfloat16_rn(s);
    }
    __syncthreads();
    HG_MARK(7);
    // ================= dec[2]: cat(e2, up(o3)) 24 -> 8 on 8x8 : 4 row-pair tiles x 18 k-steps split over 4 warp groups
    {
      const int mt = warp & 3, grp = warp >> 2;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int vy0 = 2 * mt + (lj & 1);
      for (int ks = grp * 5; ks < min(18, grp * 5 + 5); ++ks) {
        const int tap = ks >> 1, kc = ks & 1, ky = tap / 3, kx = tap - ky * 3;
        const int sy = (vy0 + ky + 1) >> 1, sx = (lr + kx + 1) >> 1;
        uint32_t addr;
        if (kc == 0 && chunk == 0) addr = smb + tE2 + (uint32_t)(((vy0 + ky) * P3 + lr + kx) * 16);
        else addr = smb + tO3 + (uint32_t)(kc * PB4 + (sy * P4 + sx) * 16);
        uint32_t a[4];
        ldsm4(a, addr);
        const uint2 w = sWf[(F_D2 + ks) * 32 + lane];
        mma_bf16(acc, a, w.x, w.y);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) sScr[2048 + ((grp * 4 + mt) * 16 + g + 8 * (q >> 1)) * 8 + 2 * t + (q & 1)] = acc[q];
    }
    __syncthreads();
    {
      const int mt = tid >> 7, r = (tid >> 3) & 15, co = tid & 7;
      float s = sBias[bD2 + co];
#pragma unroll
      for (int grp = 0; grp < 4; ++grp) s += sScr[2048 + ((grp * 4 + mt) * 16 + r) * 8 + co];
      const int y = 2 * mt + (r >> 3), x = r & 7;
      sO2[((y + 1) * P3 + x + 1) * 8 + co] = __float2bfloat16_rn(s);
    }
    __syncthreads();
    HG_MARK(8);
    // ================= dec[1]: cat(e1, up(o2)) 16 -> 8 on 16x16 : 8 warps x 2 rows, one k16 step per filter tap
    if (warp < 8) {
      const int r0 = warp * 2;
      uint2 w[3][3][1];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3][0] = sWf[(F_D1 + s) * 32 + lane];
      const float bias0 = sBias[bD1 + 2 * t], bias1 = sBias[bD1 + 2 * t + 1];
      slide_bf<2, 3, 1>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
            const int sy = (r0 + i + 1) >> 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              ldsm4(a[kx], chunk == 0 ? smb + tE1 + (uint32_t)(((r0 + i) * P2 + pixoff + kx) * 16)
                                      : smb + tO2 + (uint32_t)((sy * P3 + ((pixoff + kx + 1) >> 1)) * 16));
          },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                *reinterpret_cast<uint32_t*>(smraw + tO1 + ((r0 + e + r + 1) * P2 + g + 8 * h + 1) * 16 + 4 * t) =
                    pack_bf16((r ? bot[2 * h] : top[2 * h]) + bias0, (r ? bot[2 * h + 1] : top[2 * h + 1]) + bias1);
          });
    }
    __syncthreads();
    HG_MARK(9);
    // ================= dec[0]: cat(e0, up(o1)) 16 -> 8 on 32x32 -> o0 : 2 strips x 8 segments of 4 rows
    {
      const int x0 = (warp & 1) * 16, r0 = (warp >> 1) * 4;
      uint2 w[3][3][1];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3][0] = sWf[(F_D0 + s) * 32 + lane];
      const float bias0 = sBias[bD0 + 2 * t], bias1 = sBias[bD0 + 2 * t + 1];
      slide_bf<4, 3, 1>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
            const int sy = (r0 + i + 1) >> 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              ldsm4(a[kx], chunk == 0 ? smb + tE0 + (uint32_t)(((r0 + i) * P1 + x0 + pixoff + kx) * 16)
                                      : smb + tO1 + (uint32_t)((sy * P2 + ((x0 + pixoff + kx + 1) >> 1)) * 16));
          },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                *reinterpret_cast<uint32_t*>(smraw + tO0 + ((r0 + e + r + 1) * P1 + x0 + g + 8 * h + 1) * 16 + 4 * t) =
                    pack_bf16((r ? bot[2 * h] : top[2 * h]) + bias0, (r ? bot[2 * h + 1] : top[2 * h + 1]) + bias1);
          });
    }
    if (p.tape) fence_proxy_async();                   // this thread's plane stores become visible to the async proxy ...
    __syncthreads();
    HG_MARK(10);
    // ================= training: the tape (every plane above, exactly as it lies in shared memory) -> HBM as ONE TMA bulk
    // store (54 KB, cp.async.bulk); it drains while the masker bands run and is waited for before the next frame's first
    // write into these planes
    if (p.tape && tid == 0) bulk_s2g(p.tape + (size_t)n * TAPE, smb, TAPE);
    HG_MARK(11);
    // ================= masker.0 + LeakyReLU -> 10-row band -> masker.2 + Sigmoid (+ threshold), 8 bands of 8 mask rows.
    // masker.2 has ONE output channel: as an implicit GEMM it would fill 1 of 8 MMA columns.  Instead the MMA computes the
    // nine per-tap partial products P[pixel][tap] = sum_ci m0[pixel][ci] * W2[ci][tap] of every band pixel (K = 16 channels,
    // N = 9 taps: two MMAs per 16 pixels), and the 3x3 stencil sum z = sum_tap P[y + ky][x + kx][tap] runs with every thread
    // a pixel: 9 conflict-free shared loads, sigmoid, fully coalesced stores.
    // (the split-K scratch aliases the band: its columns 0 and 65 must be zero again)
    for (int e = tid; e < 2 * 10 * 2; e += NT) {
      const int pl = e / 20, r = (e % 20) >> 1, c = (e & 1) ? 65 : 0;
      *reinterpret_cast<uint4*>(smraw + fBand + pl * PLB8 + (r * PX + c) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    float* sP = reinterpret_cast<float*>(smraw + fP);
    for (int band = 0; band < 8; ++band) {
      if (band < 2) HG_MARK(12 + 3 * band);             // 12: band 0 starts, 15: band 1 starts
      m0_band8(smraw + fBand, smb, fX, 0, 66, sWf + F_M0 * 32, sBias + bM0, band, warp, lane);
      __syncthreads();
      if (band == 0) HG_MARK(13);
      for (int tile = warp; tile < 40; tile += 16) {
        const int r = tile >> 2, s = tile & 3;
        uint32_t a[4];
        ldsm4(a, smb + fBand + (uint32_t)(chunk * PLB8 + (r * PX + 1 + 16 * s + pixoff) * 16));
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
        const uint2 w0 = sWf[(F_SMEM_STEPS + 0) * 32 + lane], w1 = sWf[(F_SMEM_STEPS + 1) * 32 + lane];
        mma_bf16(c0, a, w0.x, w0.y);
        mma_bf16(c1, a, w1.x, w1.y);
        float* q = sP + (r * 9 + 2 * t) * PXP + 1 + 16 * s + g;
        q[0] = c0[0]; q[PXP] = c0[1]; q[8] = c0[2]; q[PXP + 8] = c0[3];
        if (t == 0) { float* q8 = sP + (r * 9 + 8) * PXP + 1 + 16 * s + g; q8[0] = c1[0]; q8[8] = c1[2]; }
      }
      __syncthreads();
      if (band == 0) HG_MARK(14);
      {
        const int o = tid >> 6, x = tid & 63;
        float acc = sBias[bM2];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc += sP[((o + ky) * 9 + ky * 3 + kx) * PXP + x + kx];
        const float zz = sigmoidf_(acc);
        const size_t off = (size_t)n * 4096 + (8 * band + o) * 64 + x;
        p.z[off] = zz;
        if (p.hard) p.hard[off] = zz >= p.thresh;
      }
    }
    HG_MARK(20);
    ++fr;
  }
  cp_async_wait_all();
  if (p.tape && tid == 0) bulk_wait_read();
  if (p.rng_state && tid == 0) {                     // last CTA to finish advances the call counter (every CTA has read it)
    __threadfence();
    if (atomicAdd(&p.rng_state[1], 1ull) == gridDim.x - 1) {
      p.rng_state[1] = 0;
      p.rng_state[0] = rng_call + 1;
    }
  }
}

}  // namespace hg
}  // namespace cgs

using namespace cgs;

namespace cgs { namespace hg { int set_fwd_trace(long long* b) { return cudaMemcpyToSymbol(g_hgf_trace, &b, sizeof(b)) == cudaSuccess ? 0 : -2; } } }

extern "C" int cgs_hg_pack_words(void) { return hg::NSTEPS * 64; }
extern "C" int cgs_hg_tape_bytes(void) { return hg::TAPE; }
extern "C" int cgs_hg_partial_stride(void) { return hg::PSTRIDE_M; }

extern "C" int cgs_hg_pack(const cgs_critic_weights* cw, const cgs_masker_weights* mw, uint32_t* pack, void* stream) {
  CGS_REQUIRE(cw && mw && pack, "hg_pack: bad args");
  CGS_REQUIRE(((uintptr_t)pack & 15) == 0, "hg_pack: pack must be 16-byte aligned");
  hg::PackSrc s;
  s.w0 = cw->w0; s.w1 = cw->w1; s.w2 = cw->w2; s.w3 = cw->w3;
  s.d0 = mw->wd0; s.d1 = mw->wd1; s.d2 = mw->wd2; s.d3 = mw->wd3; s.m0 = mw->wm0; s.m2 = mw->wm2;
  CGS_REQUIRE(s.w0 && s.w1 && s.w2 && s.w3 && s.d0 && s.d1 && s.d2 && s.d3 && s.m0 && s.m2, "hg_pack: NULL weight tensor");
  hg::hg_pack_kernel<<<(hg::NSTEPS * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(s, reinterpret_cast<uint2*>(pack));
  return check_launch("hg_pack");
}

extern "C" int cgs_hg_forward(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev,
                              const cgs_critic_weights* cw, const cgs_masker_weights* mw, const uint32_t* pack, int32_t train,
                              const float* m_e2, const float* m_e3, const float* m_v, float p_drop, uint64_t seed,
                              uint64_t* rng_state, float thresh, float* pred, float* z, uint8_t* hard, void* tape, void* stream) {
  CGS_REQUIRE(frames && cw && mw && pack && pred && z && B > 0, "hg_forward: bad args");
  CGS_REQUIRE((((uintptr_t)frames | (uintptr_t)pack | (uintptr_t)tape) & 15) == 0, "hg_forward: frames, pack and tape must be 16-byte aligned");
  CGS_REQUIRE((m_e2 != nullptr) == (m_e3 != nullptr) && (m_e2 != nullptr) == (m_v != nullptr), "hg_forward: dropout masks are all-or-none");
  CGS_REQUIRE((((uintptr_t)m_e2 | (uintptr_t)m_e3 | (uintptr_t)m_v) & 15) == 0, "hg_forward: masks must be 16-byte aligned");
  CGS_REQUIRE(!(rng_state && m_e2), "hg_forward: pass dropout masks OR an rng state, not both");
  CGS_REQUIRE(!rng_state || (p_drop > 0.f && p_drop < 1.f), "hg_forward: rng dropout needs 0 < p < 1");
  CGS_REQUIRE(train || (!m_e2 && !rng_state), "hg_forward: dropout only in train mode");
  hg::FwdParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames; p.pack = reinterpret_cast<const uint2*>(pack);
  p.b0 = cw->b0; p.b1 = cw->b1; p.b2 = cw->b2; p.b3 = cw->b3; p.w4 = cw->w4; p.b4 = cw->b4;
  p.wl1 = cw->wl1; p.bl1 = cw->bl1; p.wl2 = cw->wl2; p.bl2 = cw->bl2;
  p.wd4 = mw->wd4; p.bd4 = mw->bd4; p.bd3 = mw->bd3; p.bd2 = mw->bd2; p.bd1 = mw->bd1; p.bd0 = mw->bd0; p.bm0 = mw->bm0; p.bm2 = mw->bm2;
  CGS_REQUIRE(p.b0 && p.b1 && p.b2 && p.b3 && p.w4 && p.b4 && p.wl1 && p.bl1 && p.wl2 && p.bl2 && p.wd4 && p.bd4 && p.bd3 && p.bd2 &&
                  p.bd1 && p.bd0 && p.bm0 && p.bm2, "hg_forward: NULL parameter tensor");
  p.m2 = m_e2; p.m3 = m_e3; p.mv = m_v; p.seed = seed; p.rng_state = (unsigned long long*)rng_state;
  p.p_drop = p_drop; p.keep = rng_state ? 1.f / (1.f - p_drop) : 1.f;
  p.roll_dev = roll_dev; p.B = B; p.roll = roll; p.train = train; p.thresh = thresh;
  p.pred = pred; p.z = z; p.hard = hard; p.tape = (uint8_t*)tape;
  cudaFuncSetAttribute(hg::hg_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::F_SMEM);
  const int sms = device_sms(), per = (B + sms - 1) / sms, grid = (B + per - 1) / per;
  hg::hg_forward_kernel<<<grid, hg::NT, hg::F_SMEM, (cudaStream_t)stream>>>(p);
  return check_launch("hg_forward");
}
