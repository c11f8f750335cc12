// One critic_pipe training step (reference main.py:185-198) as ONE persistent kernel with bf16 tensor-core operands (fp32
// accumulation), chfak = 1: the bf16 successor of critic_fused.cu's TF32 kernel, built from the Hourglass kernels' blocks
// (hg_common.cuh): uint8 frame -> /255 + shift_batch roll -> NewCritic forward (dropout) -> MSE / BCE -> backward -> parameter
// gradients [-> grid barrier -> slice sum -> all-reduce over NVLink peer memory -> Adam, shared with the TF32 kernel:
// critic_tail.cuh].
//   * forward and input gradients: exactly hg_score.cu's (tap-paired k16 MMAs, pair-duplicated frame, first-max arg-max bytes,
//     pool / ReLU / dropout backward as scatters);
//   * weight gradients: GEMMs with K = pixels whose operands come from ldmatrix.trans (hg_backward.cu's sliding triples);
//     for features.0 the 64x64x8 output gradient is never materialised: the B fragment is selected on the fly from the
//     pooled-resolution gradient and the arg-max bytes (one value + one byte per 2x2 window), and the frame's fourth channel is
//     set to one so that the bias gradient rides in the same MMAs;
//   * accumulators live in registers / shared memory over all frames of the CTA; the gradient leaves as ONE partial vector per
//     CTA in the layout of critic_tail.cuh (identical to the TF32 kernel's, so everything downstream is shared).
#include <string.h>
#include "critic_tail.cuh"
#include "hg_common.cuh"

namespace cgs {
namespace hc {
using namespace hg;
using cf::Params;
using cf::aW0; using cf::aB0; using cf::aW1; using cf::aB1; using cf::aW2; using cf::aB2; using cf::aW3; using cf::aB3; using cf::aB4;
using cf::aWl1; using cf::aBl1; using cf::aWl2; using cf::aBl2; using cf::szAcc; using cf::NGRAD; using cf::PSTRIDE;

// ---- shared memory map (byte offsets)
constexpr int sX = 0;                               // pair-duplicated frame (fourth channel = 1 inside the frame)
constexpr int sE0 = sX + PBX, sDY1 = sE0 + PB1;     // e0 | d(features.3 output) (32x32x8)
constexpr int sDE0 = sDY1 + PB1;                    // d e0 at pooled resolution, dense [32*32][8] bf16
constexpr int sE1 = sDE0 + 16384, sDY2 = sE1 + PB2;
constexpr int sE2 = sDY2 + PB2;                     // e2 * dropout mask
constexpr int sDY3 = sE2 + PB3;                     // d(features.10 output) (8x8x16): 2 planes
constexpr int sI0 = sDY3 + 2 * PB3;                 // arg-max bytes [32*32][8], [16*16][8], [8*8][8], [4*4][16]
constexpr int sI1 = sI0 + 8192, sI2 = sI1 + 2048, sI3 = sI2 + 512;
constexpr int sU8 = sI3 + 256;                      // raw frame bytes (prefetched)
constexpr int sX3 = sU8 + 12288;                    // e3 * mask in the 4x4 conv's K order, fp32 [256]
constexpr int sVec = sX3 + 1024;                    // fp32: h[32] v[32] dh[32] dv[32]
constexpr int sM2 = sVec + 512, sM3 = sM2 + 2048, sMV = sM3 + 1024;
constexpr int sAccB = sMV + 128;                    // gradient accumulators (critic_tail.cuh layout), fp32 [szAcc]
constexpr int sW = sAccB + szAcc * 4;               // weight fragments: forward steps [0, 25) | input-gradient steps [B_C3D, B_C0D)
constexpr int W_BWD = F_D2, W_STEPS = F_D2 + (B_C0D - B_C3D);
constexpr int sBias = sW + W_STEPS * 256;           // b0[8] b1[8] b2[8] b3[16]
constexpr int sHW = sBias + 256;                    // wl1[1024] bl1[32] wl2[32] bl2[4] b4[32]
constexpr int hWl1 = 0, hBl1 = 1024, hWl2 = 1056, hBl2 = 1088, hB4 = 1092, szHW = 1124;
constexpr int C_SMEM = sHW + szHW * 4;
static_assert(C_SMEM <= 227 * 1024, "critic kernel: shared memory budget");
static_assert(sE0 % 16 == 0 && sDE0 % 16 == 0 && sDY3 % 16 == 0 && sU8 % 16 == 0 && sX3 % 16 == 0 && sM2 % 16 == 0 && sAccB % 16 == 0 &&
                  sW % 16 == 0, "alignment");

__device__ long long* g_hgc_trace = nullptr;
#define HC_MARK(k)                                                    \
  do {                                                                \
    if (trace && tid == 0 && fr < 2) trace[fr * 32 + (k)] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(NT, 1) hg_critic_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, odd = g & 1;
  const int lj = lane >> 3, lr = lane & 7, pixoff = lr + 8 * (lj & 1), chunk = lj >> 1;
  const int tsel = lj & 1, tpix = lr + 8 * (lj >> 1);        // ldmatrix.trans roles: matrix (lj & 1) = tap selector, (lj >> 1) = pixel half
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(smraw);
  const uint2* sWf = reinterpret_cast<const uint2*>(smraw + sW);
  float* sBi = reinterpret_cast<float*>(smraw + sBias);
  float* sHWf = reinterpret_cast<float*>(smraw + sHW);
  float* sH = reinterpret_cast<float*>(smraw + sVec);
  float *sV = sH + 32, *sDH = sH + 64, *sDV = sH + 96;
  float* fX3 = reinterpret_cast<float*>(smraw + sX3);
  float* fM2 = reinterpret_cast<float*>(smraw + sM2);
  float* fM3 = reinterpret_cast<float*>(smraw + sM3);
  float* fMV = reinterpret_cast<float*>(smraw + sMV);
  float* sAcc = reinterpret_cast<float*>(smraw + sAccB);
  uint8_t *bI0 = smraw + sI0, *bI1 = smraw + sI1, *bI2 = smraw + sI2, *bI3 = smraw + sI3;
  __nv_bfloat16* hE0 = reinterpret_cast<__nv_bfloat16*>(smraw + sE0);
  __nv_bfloat16* hE1 = reinterpret_cast<__nv_bfloat16*>(smraw + sE1);
  __nv_bfloat16* hE2 = reinterpret_cast<__nv_bfloat16*>(smraw + sE2);
  __nv_bfloat16* hDY1 = reinterpret_cast<__nv_bfloat16*>(smraw + sDY1);
  __nv_bfloat16* hDY2 = reinterpret_cast<__nv_bfloat16*>(smraw + sDY2);
  __nv_bfloat16* hDY3 = reinterpret_cast<__nv_bfloat16*>(smraw + sDY3);
  const unsigned short* uDE0 = reinterpret_cast<const unsigned short*>(smraw + sDE0);
  long long* trace = blockIdx.x == 0 ? g_hgc_trace : nullptr;
  int fr = 0;
  const uint32_t ones = g == 0 ? 0x3F803F80u : 0u;

  const unsigned long long rng_call = p.rng_state ? p.rng_state[0] : 0ull;
  unsigned bar_gen = 0;
  if (p.adam_p && tid == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(bar_gen) : "l"(p.bar + 1) : "memory");
  if (blockIdx.x < p.B) {
    const uint8_t* src = p.frames + (size_t)blockIdx.x * 12288;
    for (int c = tid; c < 768; c += NT) cp_async16(smb + sU8 + c * 16, src + c * 16);
    cp_async_commit();
  }
  // ---- prologue: zero every plane and the accumulators, weight fragments (packed by cgs_hg_pack), biases, head weights
  {
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int e = tid; e < sI0 / 16; e += NT) reinterpret_cast<uint4*>(smraw)[e] = z4;
    for (int e = tid; e < szAcc; e += NT) sAcc[e] = 0.f;
    // weight fragments straight from the OIHW parameters (they change every step: no pack launch on this path); all loads of
    // a thread are independent, so a cold launch pays one L2 / HBM round trip for them
    PackSrc ps;
    ps.w0 = p.w0; ps.w1 = p.w1; ps.w2 = p.w2; ps.w3 = p.w3;
    ps.d0 = ps.d1 = ps.d2 = ps.d3 = ps.m0 = ps.m2 = nullptr;
    for (int e = tid; e < W_STEPS * 32; e += NT) {
      const int sl = e >> 5, ln = e & 31, gg = ln >> 2, tt = ln & 3, s = sl < F_D2 ? sl : B_C3D + (sl - F_D2);
      uint2 v;
      v.x = pack_bf16(pack_wk(ps, s, 2 * tt, gg), pack_wk(ps, s, 2 * tt + 1, gg));
      v.y = pack_bf16(pack_wk(ps, s, 2 * tt + 8, gg), pack_wk(ps, s, 2 * tt + 9, gg));
      reinterpret_cast<uint2*>(smraw + sW)[e] = v;
    }
  }
  if (tid < 8) { sBi[tid] = __ldg(p.b0 + tid); sBi[8 + tid] = __ldg(p.b1 + tid); sBi[16 + tid] = __ldg(p.b2 + tid); }
  if (tid < 16) sBi[24 + tid] = __ldg(p.b3 + tid);
  for (int e = tid; e < 1024; e += NT) sHWf[hWl1 + e] = __ldg(p.wl1 + e);
  if (tid < 32) { sHWf[hBl1 + tid] = __ldg(p.bl1 + tid); sHWf[hWl2 + tid] = __ldg(p.wl2 + tid); sHWf[hB4 + tid] = __ldg(p.b4 + tid); }
  if (tid == 0) sHWf[hBl2] = __ldg(p.bl2);
  if (!p.m2 && !p.rng_state) {                         // eval mode / p = 0: identity masks, written once
    fM2[tid] = 1.f;
    if (tid < 256) fM3[tid] = 1.f;
    if (tid < 32) fMV[tid] = 1.f;
  }
  int roll = p.roll_dev ? *p.roll_dev : p.roll;
  roll = ((roll % 64) + 64) & 63;
  // gradient accumulators in registers over all frames of this CTA: features.14 (16 per thread), features.0 (every warp a
  // partial), features.3 (warps 0-7) / features.6 (warps 8-15) triples, features.10 (warps 0-9, one tile each)
  float accW4[16], acc0[3][4], accW[3][4], acc3[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 16; ++i) accW4[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) { acc0[i >> 2][i & 3] = 0.f; accW[i >> 2][i & 3] = 0.f; }
  float loss_acc = 0.f;

  for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
    cp_async_wait_all();
    __syncthreads();
    HC_MARK(0);
    // ================= frame bytes have landed -> pair-duplicated bf16 plane, fourth channel = 1; dropout masks of this frame
    for (int e = tid; e < 4096; e += NT) {
      const int y = e >> 6, x = e & 63;
      const uint8_t* s = smraw + sU8 + (y * 64 + ((x + roll) & 63)) * 3;
      constexpr float k = 1.f / 255.f;                 // bf16(b * fl(1/255)) == bf16(b / 255.0f) for all 256 bytes
      const uint2 q = make_uint2(pack_bf16(__fmul_rn((float)s[0], k), __fmul_rn((float)s[1], k)), pack_bf16(__fmul_rn((float)s[2], k), 1.f));
      uint8_t* row = smraw + sX + (size_t)(y + 1) * (PX * 16);
      *reinterpret_cast<uint2*>(row + (x + 1) * 16) = q;
      *reinterpret_cast<uint2*>(row + x * 16 + 8) = q;
    }
    if (p.rng_state) draw_masks3(p.seed, rng_call, p.B, n, p.p_drop, p.keep, fM2, fM3, fMV, tid);
    else if (p.m2) {
      if (tid < 128) reinterpret_cast<float4*>(fM2)[tid] = __ldg(reinterpret_cast<const float4*>(p.m2 + (size_t)n * 512) + tid);
      else if (tid < 192) reinterpret_cast<float4*>(fM3)[tid - 128] = __ldg(reinterpret_cast<const float4*>(p.m3 + (size_t)n * 256) + tid - 128);
      else if (tid < 200) reinterpret_cast<float4*>(fMV)[tid - 192] = __ldg(reinterpret_cast<const float4*>(p.mv + (size_t)n * 32) + tid - 192);
    }
    const float ytgt = __ldg(p.target + n);
    __syncthreads();
    if (n + (int)gridDim.x < p.B) {
      const uint8_t* src = p.frames + (size_t)(n + gridDim.x) * 12288;
      for (int c = tid; c < 768; c += NT) cp_async16(smb + sU8 + c * 16, src + c * 16);
      cp_async_commit();
    }
    HC_MARK(1);
    // ================= F0: features.0 (3 -> 8) + ReLU + pool + arg-max -> e0
    {
      const int x0 = (warp & 3) * 16, r0 = (warp >> 2) * 16;
      uint2 w[3][1][1];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) w[ky][0][0] = sWf[(F_C0 + ky) * 32 + lane];
      const uint32_t aA = smb + sX + (uint32_t)((r0 * PX + x0 + pixoff + 2 * chunk) * 16);
      const float bias0 = sBi[2 * t], bias1 = sBi[2 * t + 1];
      const int co = 2 * t + odd;
      __nv_bfloat16* dE = hE0 + (((r0 >> 1) + 1) * P1 + ((x0 + g) >> 1) + 1) * 8 + co;
      uint8_t* dI = bI0 + ((r0 >> 1) * 32 + ((x0 + g) >> 1)) * 8 + co;
      slide_bf<16, 1, 1>(
          w, [&](int i, uint32_t(&a)[1][4]) { ldsm4(a[0], aA + (uint32_t)(i * (PX * 16))); },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
            pool2x2(top, bot, bias0, bias1, odd, [&](int h, float v, int idx) {
              dE[((e >> 1) * P1 + 4 * h) * 8] = __float2bfloat16_rn(v);
              dI[((e >> 1) * 32 + 4 * h) * 8] = (uint8_t)idx;
            });
          });
    }
    __syncthreads();
    HC_MARK(2);
    // ================= F1: features.3 (8 -> 8) on 32x32 -> e1; the scatter targets of the backward are cleared meanwhile
    {
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      for (int e = tid; e < PB1 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sDY1)[e] = z4;
      for (int e = tid; e < PB2 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sDY2)[e] = z4;
      for (int e = tid; e < 2 * PB3 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sDY3)[e] = z4;
      const int x0 = (warp & 1) * 16, r0 = (warp >> 1) * 4;
      uint2 w[3][2][1];
#pragma unroll
      for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(F_C1 + s) * 32 + lane];
      const uint32_t aA = smb + sE0 + (uint32_t)((r0 * P1 + x0 + pixoff + chunk) * 16);
      const uint32_t aB = smb + sE0 + (uint32_t)((r0 * P1 + x0 + pixoff + 2) * 16);
      const float bias0 = sBi[8 + 2 * t], bias1 = sBi[8 + 2 * t + 1];
      const int co = 2 * t + odd;
      __nv_bfloat16* dE = hE1 + (((r0 >> 1) + 1) * P2 + ((x0 + g) >> 1) + 1) * 8 + co;
      uint8_t* dI = bI1 + ((r0 >> 1) * 16 + ((x0 + g) >> 1)) * 8 + co;
      slide_bf<4, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P1 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P1 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
            pool2x2(top, bot, bias0, bias1, odd, [&](int h, float v, int idx) {
              dE[((e >> 1) * P2 + 4 * h) * 8] = __float2bfloat16_rn(v);
              dI[((e >> 1) * 16 + 4 * h) * 8] = (uint8_t)idx;
            });
          });
    }
    __syncthreads();
    HC_MARK(3);
    // ================= F2: features.6 (8 -> 8) on 16x16 + Dropout -> e2 * mask
    if (warp < 8) {
      const int r0 = warp * 2;
      uint2 w[3][2][1];
#pragma unroll
      for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(F_C2 + s) * 32 + lane];
      const uint32_t aA = smb + sE1 + (uint32_t)((r0 * P2 + pixoff + chunk) * 16);
      const uint32_t aB = smb + sE1 + (uint32_t)((r0 * P2 + pixoff + 2) * 16);
      const float bias0 = sBi[16 + 2 * t], bias1 = sBi[16 + 2 * t + 1];
      slide_bf<2, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P2 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P2 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int, int, const float(&top)[4], const float(&bot)[4]) {
            pool2x2(top, bot, bias0, bias1, odd, [&](int h, float v, int idx) {
              const int py = warp, px = (g >> 1) + 4 * h, co = 2 * t + odd, q = (py * 8 + px) * 8 + co;
              hE2[((py + 1) * P3 + px + 1) * 8 + co] = __float2bfloat16_rn(v * fM2[q]);
              bI2[q] = (uint8_t)idx;
            });
          });
    }
    __syncthreads();
    HC_MARK(4);
    // ================= F3: features.10 (8 -> 16) on 8x8 + Dropout -> head operand (K order) + arg-max
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
        ldsm4(a, smb + sE2 + (uint32_t)(((2 * mt + (lj & 1) + ky) * P3 + lr + kx) * 16));
        const uint2 w = sWf[(F_C3 + tp * 2 + nt) * 32 + lane];
        mma_bf16(acc, a, w.x, w.y);
      }
      const int co = nt * 8 + 2 * t + odd;
      const float b0 = sBi[24 + nt * 8 + 2 * t], b1 = sBi[24 + nt * 8 + 2 * t + 1];
      const float t0 = acc[0] + b0, t1 = acc[1] + b1, u0 = acc[2] + b0, u1 = acc[3] + b1;
      const float rt = __shfl_xor_sync(0xffffffffu, odd ? t0 : t1, 4), rb = __shfl_xor_sync(0xffffffffu, odd ? u0 : u1, 4);
      const float p0 = odd ? rt : t0, p1 = odd ? t1 : rt, p2 = odd ? rb : u0, p3 = odd ? u1 : rb;
      const float m01 = fmaxf(p0, p1), m23 = fmaxf(p2, p3);
      const int i01 = p1 > p0 ? 1 : 0, i23 = p3 > p2 ? 3 : 2;
      float m = fmaxf(m01, m23);
      int idx = m23 > m01 ? i23 : i01;
      if (!(m > 0.f)) { m = 0.f; idx = 4; }
      const int pp = mt * 4 + (g >> 1);
      fX3[co * 16 + pp] = m * fM3[pp * 16 + co];
      bI3[pp * 16 + co] = (uint8_t)idx;
    }
    __syncthreads();
    HC_MARK(5);
    // ================= F4: features.14 (4x4 valid conv = 256 -> 32) + ReLU
    {
      const int nn = tid >> 4, part = tid & 15;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 aq = w4r[i];
        const float4 bq = *reinterpret_cast<const float4*>(fX3 + part * 16 + ((i + rot4) & 3) * 4);
        s = fmaf(aq.x, bq.x, s); s = fmaf(aq.y, bq.y, s); s = fmaf(aq.z, bq.z, s); s = fmaf(aq.w, bq.w, s);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0) sH[nn] = fmaxf(s + sHWf[hB4 + nn], 0.f);
    }
    __syncthreads();
    // ================= F5: crit.1 Linear(32,32) + ReLU
    {
      const int nn = tid >> 4, part = tid & 15;
      const float2 wv = *reinterpret_cast<const float2*>(sHWf + hWl1 + nn * 32 + 2 * part);
      float s = wv.x * sH[2 * part] + wv.y * sH[2 * part + 1];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0) sV[nn] = fmaxf(s + sHWf[hBl1 + nn], 0.f);
    }
    __syncthreads();
    // ================= F6: Dropout, crit.4 Linear(32,1), Sigmoid, loss and its gradient; head weight gradients
    float w4c[16];
    {
      const float* wc = p.w4 + ((tid & 1) * 16) * 256 + (tid >> 1);
#pragma unroll
      for (int i = 0; i < 16; ++i) w4c[i] = __ldg(wc + i * 256);
    }
    if (warp == 0) {
      const float wk = sHWf[hWl2 + lane], vm = sV[lane] * fMV[lane];
      const float zz = warp_sum(wk * vm) + sHWf[hBl2];
      const float pr = sigmoidf_(zz), y = ytgt;
      float dl;
      if (p.bce) {
        loss_acc -= y * fmaxf(logf(pr), -100.f) + (1.f - y) * fmaxf(logf(1.f - pr), -100.f);
        dl = p.gscale * (pr - y) / fmaxf(pr * (1.f - pr), 1e-12f) * pr * (1.f - pr);
      } else {
        loss_acc = fmaf(pr - y, pr - y, loss_acc);
        dl = p.gscale * 2.f * (pr - y) * pr * (1.f - pr);
      }
      if (lane == 0) { p.pred[n] = pr; sAcc[aBl2] += dl; }
      sAcc[aWl2 + lane] += dl * vm;
      sDV[lane] = sV[lane] > 0.f ? dl * wk * fMV[lane] : 0.f;
    }
    __syncthreads();
    HC_MARK(6);
    // ================= B5: crit.1 backward
    {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = tid + NT * i;
        sAcc[aWl1 + e] = fmaf(sDV[e >> 5], sH[e & 31], sAcc[aWl1 + e]);
      }
      if (tid < 32) sAcc[aBl1 + tid] += sDV[tid];
      const int k = tid >> 4, part = tid & 15;
      float s = sHWf[hWl1 + (2 * part) * 32 + k] * sDV[2 * part] + sHWf[hWl1 + (2 * part + 1) * 32 + k] * sDV[2 * part + 1];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0) sDH[k] = sH[k] > 0.f ? s : 0.f;
    }
    __syncthreads();
    // ================= B4: features.14 backward (weight gradient in registers), Dropout + pool + ReLU backward -> d(features.10 output)
    {
      const int nn = tid >> 4, part = tid & 15;
      const float d = sDH[nn];
#pragma unroll
      for (int i = 0; i < 4; ++i) {                                         // accW4[4i..] <-> chunk (i + rot4) & 3, as in F4
        const float4 x = *reinterpret_cast<const float4*>(fX3 + part * 16 + ((i + rot4) & 3) * 4);
        accW4[4 * i + 0] = fmaf(d, x.x, accW4[4 * i + 0]); accW4[4 * i + 1] = fmaf(d, x.y, accW4[4 * i + 1]);
        accW4[4 * i + 2] = fmaf(d, x.z, accW4[4 * i + 2]); accW4[4 * i + 3] = fmaf(d, x.w, accW4[4 * i + 3]);
      }
      if (tid < 32) sAcc[aB4 + tid] += sDH[tid];
      const int k = tid >> 1, hf = tid & 1;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s = fmaf(w4c[i], sDH[hf * 16 + i], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (hf == 0) {
        const int co = k >> 4, pp = k & 15, idx = bI3[pp * 16 + co];
        if (idx < 4) {
          const int y = 2 * (pp >> 2) + (idx >> 1), x = 2 * (pp & 3) + (idx & 1);
          hDY3[(co >> 3) * (PB3 / 2) + ((y + 1) * P3 + x + 1) * 8 + (co & 7)] = __float2bfloat16_rn(s * fM3[pp * 16 + co]);
        }
      }
    }
    __syncthreads();
    HC_MARK(7);
    // ================= B3: features.10 weight gradient (warps 0-9: one (tap pair, channel tile) tile each, registers)
    //                   || input gradient (warps 10-13), Dropout + pool + ReLU backward -> d(features.6 output)
    if (warp < 10) {
      const int tp = warp >> 1, nt = warp & 1;
      const int tap = min(2 * tp + tsel, 8), ky = tap / 3, kx = tap - 3 * ky;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        uint32_t a[4];
        ldsm4t(a, smb + sE2 + (uint32_t)(((2 * mt + (lj >> 1) + ky) * P3 + lr + kx) * 16));
        if (tp == 4) a[1] = a[3] = ones;
        uint32_t b0, b1;
        ldsm2t(b0, b1, smb + sDY3 + (uint32_t)(nt * PB3 + ((2 * mt + (lj & 1) + 1) * P3 + 1 + lr) * 16));
        mma_bf16(acc3, a, b0, b1);
      }
    } else if (warp < 14) {
      const int mt = warp - 10;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const int ky = tp / 3, kx = tp - 3 * ky;
        uint32_t a[4];
        ldsm4(a, smb + sDY3 + (uint32_t)(chunk * PB3 + ((2 * mt + (lj & 1) + ky) * P3 + lr + kx) * 16));
        const uint2 w = sWf[(W_BWD + tp) * 32 + lane];
        mma_bf16(acc, a, w.x, w.y);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int y = 2 * mt + (q >> 1), ci = 2 * t + (q & 1), pq = (y * 8 + g) * 8 + ci, idx = bI2[pq];
        if (idx < 4)
          hDY2[((2 * y + (idx >> 1) + 1) * P2 + 2 * g + (idx & 1) + 1) * 8 + ci] = __float2bfloat16_rn(acc[q] * fM2[pq]);
      }
    }
    __syncthreads();
    HC_MARK(8);
    // ================= B2: features.6 input gradient (warps 0-7) -> d(features.3 output) || weight gradient (warps 8-15, registers)
    if (warp < 8) {
      const int r0 = warp * 2;
      uint2 w[3][2][1];
#pragma unroll
      for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(W_BWD + B_C2D - B_C3D + s) * 32 + lane];
      const uint32_t aA = smb + sDY2 + (uint32_t)((r0 * P2 + pixoff + chunk) * 16);
      const uint32_t aB = smb + sDY2 + (uint32_t)((r0 * P2 + pixoff + 2) * 16);
      slide_bf<2, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P2 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P2 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int y = r0 + e + r, x = g + 8 * (q >> 1), ci = 2 * t + (q & 1), idx = bI1[(y * 16 + x) * 8 + ci];
                if (idx < 4)
                  hDY1[((2 * y + (idx >> 1) + 1) * P1 + 2 * x + (idx & 1) + 1) * 8 + ci] = __float2bfloat16_rn(r ? bot[q] : top[q]);
              }
          });
    } else {
      // features.6: input e1 (16x16), gradient dY2; triple = kx group (warp & 1), rows [4q, 4q+4) with q = (warp - 8) >> 1
      const int kxg = warp & 1, y0 = ((warp - 8) >> 1) * 4;
      auto loadB = [&](int y, uint32_t& b0, uint32_t& b1) {
        ldsm2t(b0, b1, smb + sDY2 + (uint32_t)(((y + 1) * P2 + 1 + lr + 8 * (lj & 1)) * 16));
      };
      if (kxg == 0)
        wgrad_slide<4>(accW, y0, [&](int i, uint32_t(&a)[4]) { ldsm4t(a, smb + sE1 + (uint32_t)((i * P2 + tpix + tsel) * 16)); }, loadB);
      else
        wgrad_slide<4>(accW, y0,
                     [&](int i, uint32_t(&a)[4]) {
                       ldsm2t(a[0], a[2], smb + sE1 + (uint32_t)((i * P2 + lr + 8 * (lj & 1) + 2) * 16));
                       a[1] = a[3] = ones;
                     },
                     loadB);
    }
    __syncthreads();
    HC_MARK(9);
    // ================= B1: features.3 weight gradient (warps 0-7, registers) || input gradient (warps 8-15) -> d e0 at pooled
    //                   resolution (dense: the arg-max bytes say where each value belongs at full resolution)
    if (warp < 8) {
      // input e0 (32x32), gradient dY1; triple = kx group (warp & 1), strip (warp >> 1) & 1, rows [16h, 16h+16) with h = warp >> 2
      const int kxg = warp & 1, x0 = ((warp >> 1) & 1) * 16, y0 = (warp >> 2) * 16;
      auto loadB = [&](int y, uint32_t& b0, uint32_t& b1) {
        ldsm2t(b0, b1, smb + sDY1 + (uint32_t)(((y + 1) * P1 + 1 + x0 + lr + 8 * (lj & 1)) * 16));
      };
      if (kxg == 0)
        wgrad_slide<16>(accW, y0, [&](int i, uint32_t(&a)[4]) { ldsm4t(a, smb + sE0 + (uint32_t)((i * P1 + x0 + tpix + tsel) * 16)); }, loadB);
      else
        wgrad_slide<16>(accW, y0,
                      [&](int i, uint32_t(&a)[4]) {
                        ldsm2t(a[0], a[2], smb + sE0 + (uint32_t)((i * P1 + x0 + lr + 8 * (lj & 1) + 2) * 16));
                        a[1] = a[3] = ones;
                      },
                      loadB);
    } else {
      const int x0 = (warp & 1) * 16, r0 = ((warp - 8) >> 1) * 8;
      uint2 w[3][2][1];
#pragma unroll
      for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(W_BWD + B_C1D - B_C3D + s) * 32 + lane];
      const uint32_t aA = smb + sDY1 + (uint32_t)((r0 * P1 + x0 + pixoff + chunk) * 16);
      const uint32_t aB = smb + sDY1 + (uint32_t)((r0 * P1 + x0 + pixoff + 2) * 16);
      slide_bf<8, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P1 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P1 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                *reinterpret_cast<uint32_t*>(smraw + sDE0 + ((r0 + e + r) * 32 + x0 + g + 8 * h) * 16 + 4 * t) =
                    pack_bf16(r ? bot[2 * h] : top[2 * h], r ? bot[2 * h + 1] : top[2 * h + 1]);
          });
    }
    __syncthreads();
    HC_MARK(10);
    // ================= B0: features.0 weight (+ bias) gradient: A = the pair-duplicated frame (rows = (kx, channel), channel 3
    // = 1), B = the 64x64x8 output gradient selected on the fly from pooled d e0 + arg-max: pixel x of row y gets the pooled value
    // iff the window's first max sits at (y & 1, x & 1).  16 warps = 4 strips x 4 row quarters, registers
    {
      const int x0 = (warp & 3) * 16, y0 = (warp >> 2) * 16;
      uint32_t bq[3][2];
#pragma unroll
      for (int i = 0; i < 18; ++i) {
        uint32_t a[4];
        ldsm4t(a, smb + sX + (uint32_t)(((y0 + i) * PX + x0 + tpix + 2 * tsel) * 16));
        if (i < 16) {
          const int y = y0 + i, py = y >> 1;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {            // b0: pixels x0 + 2t, 2t+1 -> pooled px (x0 >> 1) + t; b1: + 8 -> pooled px + 4
            const int o = (py * 32 + (x0 >> 1) + t + 4 * hh) * 8 + g;
            const uint32_t v = uDE0[o], idx = bI0[o];
            const uint32_t lo = idx == (uint32_t)((y & 1) << 1) ? v : 0u, hi = idx == (uint32_t)(((y & 1) << 1) | 1) ? v : 0u;
            bq[i % 3][hh] = lo | (hi << 16);
          }
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = i - ky;
          if (yy >= 0 && yy < 16) mma_bf16(acc0[ky], a, bq[yy % 3][0], bq[yy % 3][1]);
        }
      }
    }
    HC_MARK(11);
    ++fr;
    // the next iteration's first barrier separates these reads from the next frame's writes
  }

  // ---- end of the CTA's frames: all register accumulators -> the shared accumulator block (fixed order), then hand over
  cp_async_wait_all();
  __syncthreads();
  {
    float* scr = reinterpret_cast<float*>(smraw + sX);            // [16 warps][3 ky][128] features.0 | [16][3][128] conv1/2 | [10][128] features.10
    float* scrW = scr + 16 * 384;
    float* scr3 = scrW + 16 * 384;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      *reinterpret_cast<float4*>(scr + (warp * 3 + ky) * 128 + lane * 4) = make_float4(acc0[ky][0], acc0[ky][1], acc0[ky][2], acc0[ky][3]);
      *reinterpret_cast<float4*>(scrW + (warp * 3 + ky) * 128 + lane * 4) = make_float4(accW[ky][0], accW[ky][1], accW[ky][2], accW[ky][3]);
    }
    if (warp < 10) *reinterpret_cast<float4*>(scr3 + warp * 128 + lane * 4) = make_float4(acc3[0], acc3[1], acc3[2], acc3[3]);
    __syncthreads();
    // fragment value q of lane (g, t): row m = g + 8*(q >> 1), column n = 2t + (q & 1) = output channel
    for (int e = tid; e < 3 * 128; e += NT) {          // features.0 [8][3][3][3] + bias: rows m = kx*4 + c; c = 3 is the ones channel
      const int ky = e >> 7, ln = (e >> 2) & 31, q = e & 3, m = (ln >> 2) + 8 * (q >> 1), co = 2 * (ln & 3) + (q & 1), kx = m >> 2, c = m & 3;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) v += scr[(w * 3 + ky) * 128 + ln * 4 + q];
      if (kx < 3 && c < 3) sAcc[aW0 + (co * 3 + c) * 9 + ky * 3 + kx] = v;
      else if (ky == 1 && kx == 1 && c == 3) sAcc[aB0 + co] = v;    // centre tap of the ones channel: sum of the output gradient
    }
    for (int e = tid; e < 2 * 2 * 3 * 128; e += NT) {  // features.3 (warps 0-7) / features.6 (warps 8-15): [layer][kx group][ky][128]
      const int layer = e / 768, kxg = (e / 384) & 1, ky = (e >> 7) % 3, ln = (e >> 2) & 31, q = e & 3;
      const int m = (ln >> 2) + 8 * (q >> 1), co = 2 * (ln & 3) + (q & 1);
      float v = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) v += scrW[((layer * 8 + c4 * 2 + kxg) * 3 + ky) * 128 + ln * 4 + q];
      const int aw = layer ? aW2 : aW1, ab = layer ? aB2 : aB1;
      if (kxg == 0) sAcc[aw + (co * 8 + (m & 7)) * 9 + ky * 3 + (m >> 3)] = v;
      else if (m < 8) sAcc[aw + (co * 8 + m) * 9 + ky * 3 + 2] = v;
      else if (m == 8 && ky == 0) sAcc[ab + co] = v;
    }
    for (int e = tid; e < 10 * 128; e += NT) {         // features.10 [16][8][3][3] + bias: tile = tp*2 + nt
      const int tile = e >> 7, ln = (e >> 2) & 31, q = e & 3, tp = tile >> 1, nt = tile & 1;
      const int m = (ln >> 2) + 8 * (q >> 1), co = nt * 8 + 2 * (ln & 3) + (q & 1), tap = 2 * tp + (m >> 3);
      const float v = scr3[tile * 128 + ln * 4 + q];
      if (tap < 9) sAcc[aW3 + (co * 8 + (m & 7)) * 9 + tap] = v;
      else if (m == 8) sAcc[aB3 + co] = v;
    }
    __syncthreads();
  }
  cf::grad_handover(p, sAcc, accW4, tid);
  if (tid == 0) {
    if (p.adam_p) p.partials[(size_t)blockIdx.x * PSTRIDE + NGRAD] = loss_acc * p.inv_n;
    else atomicAdd(p.loss, loss_acc * p.inv_n);
  }
  if (p.adam_p) cf::adam_tail(p, reinterpret_cast<float*>(smraw + sX), bar_gen, tid, warp, lane, [&](int) {});
  if (p.rng_state && tid == 0) {                     // last CTA to finish advances the call counter (every CTA has read it)
    __threadfence();
    if (atomicAdd(&p.rng_state[1], 1ull) == gridDim.x - 1) {
      p.rng_state[1] = 0;
      p.rng_state[0] = rng_call + 1;
    }
  }
}

}  // namespace hc
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_hg_set_trace_critic(long long* buf) {
  return cudaMemcpyToSymbol(hc::g_hgc_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}

// The bf16 variant of cgs_critic_train_fused: same arguments, same outputs, same partial-vector / in-kernel Adam / peer-memory
// all-reduce contract (critic_tail.cuh).  Gradient delivery is by partial vectors only (partials != NULL).
extern "C" int cgs_critic_train_bf16(const uint8_t* frames, const float* target, int32_t B, int32_t roll, const int32_t* roll_dev,
                                     const float* m_e2, const float* m_e3, const float* m_v, float p_drop, uint64_t seed,
                                     uint64_t* rng_state, const cgs_critic_weights* w, float* partials, const cgs_adam_args* adam,
                                     float loss_grad, int32_t bce, float* pred, float* loss, void* stream) {
  CGS_REQUIRE(frames && target && w && pred && loss && partials && B > 0, "critic_train_bf16: bad args");
  CGS_REQUIRE(((uintptr_t)frames & 15) == 0, "critic_train_bf16: frames must be 16-byte aligned");
  CGS_REQUIRE((m_e2 != nullptr) == (m_e3 != nullptr) && (m_e2 != nullptr) == (m_v != nullptr), "critic_train_bf16: dropout masks are all-or-none");
  CGS_REQUIRE((((uintptr_t)m_e2 | (uintptr_t)m_e3 | (uintptr_t)m_v | (uintptr_t)partials) & 15) == 0,
              "critic_train_bf16: masks and partials must be 16-byte aligned");
  CGS_REQUIRE(!(rng_state && m_e2), "critic_train_bf16: pass dropout masks OR an rng state, not both");
  CGS_REQUIRE(!rng_state || (p_drop > 0.f && p_drop < 1.f), "critic_train_bf16: rng dropout needs 0 < p < 1");
  cudaStream_t st = (cudaStream_t)stream;
  cf::Params p;
  memset(&p, 0, sizeof(p));
  p.frames = frames; p.target = target; p.m2 = m_e2; p.m3 = m_e3; p.mv = m_v;
  p.w0 = w->w0; p.b0 = w->b0; p.w1 = w->w1; p.b1 = w->b1; p.w2 = w->w2; p.b2 = w->b2; p.w3 = w->w3; p.b3 = w->b3;
  p.w4 = w->w4; p.b4 = w->b4; p.wl1 = w->wl1; p.bl1 = w->bl1; p.wl2 = w->wl2; p.bl2 = w->bl2;
  p.seed = seed; p.rng_state = (unsigned long long*)rng_state; p.p_drop = p_drop;
  p.keep = rng_state ? 1.f / (1.f - p_drop) : 1.f;
  p.world = 1;
  const int grid = cgs_critic_fused_grid(B);
  if (adam) {
    CGS_REQUIRE(adam->p && adam->g && adam->m && adam->v && adam->step_state && adam->barrier,
                "critic_train_bf16: in-kernel Adam needs all optimizer pointers");
    CGS_REQUIRE(grid <= 152, "critic_train_bf16: grid too large for the in-kernel reduction");
    p.adam_p = adam->p; p.adam_g = adam->g; p.adam_m = adam->m; p.adam_v = adam->v;
    p.lr = adam->lr; p.beta1 = adam->beta1; p.beta2 = adam->beta2; p.eps = adam->eps;
    p.step_state = adam->step_state; p.bar = adam->barrier;
    if (adam->world > 1) {
      CGS_REQUIRE(adam->world <= 16 && adam->rank >= 0 && adam->rank < adam->world && adam->peer_recv && adam->npad >= cf::NGRAD,
                  "critic_train_bf16: bad peer-memory arguments (world %d rank %d)", adam->world, adam->rank);
      p.world = adam->world; p.rank = adam->rank; p.npad = adam->npad;
      for (int r = 0; r < adam->world; ++r) p.ll_peer[r] = reinterpret_cast<unsigned long long*>(adam->peer_recv[r]);
    }
  }
  p.partials = partials;
  p.pred = pred; p.loss = loss; p.roll_dev = roll_dev; p.B = B; p.roll = roll; p.bce = bce;
  p.inv_n = 1.f / (float)B;
  p.gscale = loss_grad / (float)B;
  cudaFuncSetAttribute(hc::hg_critic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hc::C_SMEM);
  if (!adam && cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) return check_launch("critic_train_bf16.memset");
  if (adam) {
    void* args[] = {(void*)&p};
    const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)hc::hg_critic_kernel, dim3(grid), dim3(hg::NT), args, (size_t)hc::C_SMEM, st);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      set_error("critic_train_bf16: cooperative launch of %d CTAs failed: %s", grid, cudaGetErrorString(ce));
      return CGS_ECUDA;
    }
  } else {
    hc::hg_critic_kernel<<<grid, hg::NT, hc::C_SMEM, st>>>(p);
  }
  return check_launch("critic_train_bf16");
}
