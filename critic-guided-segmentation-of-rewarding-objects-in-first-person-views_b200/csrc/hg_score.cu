// The critic-scoring part of one frozen-critic segmentation_training step (reference main.py:365-367, 395-429) as ONE
// persistent kernel with bf16 tensor-core operands (fp32 accumulation), chfak = 1.  Per frame, with everything in shared memory:
//   pass 0 (optional): negpred = critic(B)                                        main.py:365-367 (no_grad)
//   pass 1: replaced = A*(1-Z) + Z*B -> critic -> (pred - negpred)^2 -> backward to the blend -> dZ   main.py:395-400
//   pass 2: injected = B*(1-Z) + Z*A -> critic -> (pred - critic(A))^2 -> backward -> dZ              main.py:406-411
//   + the L1 / L2 mask regulariser and its gradient                                                    main.py:415-429
// Neither a blend, nor a critic activation, nor an input gradient ever exists in HBM: in = the two raw uint8 frames and
// the mask (40 KB per frame), out = d loss / d mask (16 KB) and three scalars per frame.
// Layouts and building blocks are those of hg_common.cuh (the Hourglass forward's encoder, plus first-max arg-max bytes);
// the backward (pool / ReLU / dropout backward as scatters, the four input-gradient convolutions with rotated filters)
// mirrors the TF32 kernel it replaces (critic_fused.cu, MODE 3), at half the shared-memory bytes and MMA count.
#include <string.h>
#include "hg_common.cuh"

namespace cgs {
namespace hs {
using namespace hg;

// ---- shared memory map (byte offsets)
constexpr int sX = 0;                               // pair-duplicated blend; after features.0: d(features.0 output), 64x64x8
constexpr int sE0 = sX + PBX, sDY1 = sE0 + PB1;     // e0 | d(features.3 output) (32x32x8)
constexpr int sE1 = sDY1 + PB1, sDY2 = sE1 + PB2;   // e1 | d(features.6 output) (16x16x8)
constexpr int sE2 = sDY2 + PB2;                     // e2 * dropout mask
constexpr int sDY3 = sE2 + PB3;                     // d(features.10 output) (8x8x16): 2 planes
constexpr int sI0 = sDY3 + 2 * PB3;                 // arg-max bytes [32*32][8], [16*16][8], [8*8][8], [4*4][16]
constexpr int sI1 = sI0 + 8192, sI2 = sI1 + 2048, sI3 = sI2 + 512;
constexpr int sA8 = sI3 + 256, sB8 = sA8 + 12288;   // raw frame bytes
constexpr int sZ = sB8 + 12288;                     // the frame's mask, fp32 [64][64]
constexpr int sDZ = sZ + 16384;                     // d loss / d mask accumulated over the passes, fp32 [64][64]
constexpr int sX3 = sDZ + 16384;                    // e3 * mask in the 4x4 conv's K order, fp32 [256]
constexpr int sVec = sX3 + 1024;                    // fp32: h[32] v[32] dh[32] dv[32] | scalars
constexpr int sM2 = sVec + 640, sM3 = sM2 + 2048, sMV = sM3 + 1024;
constexpr int sW = sMV + 128;                       // weight fragments: forward steps [0, 25) | input-gradient steps [B_C3D, NSTEPS)
constexpr int W_FWD = 0, W_BWD = F_D2, W_STEPS = F_D2 + (NSTEPS - B_C3D);
constexpr int sBias = sW + W_STEPS * 256;           // b0[8] b1[8] b2[8] b3[16]
constexpr int sHW = sBias + 256;                    // wl1[1024] bl1[32] wl2[32] bl2[4] b4[32]
constexpr int hWl1 = 0, hBl1 = 1024, hWl2 = 1056, hBl2 = 1088, hB4 = 1092, szHW = 1124;
constexpr int S_SMEM = sHW + szHW * 4;
static_assert(S_SMEM <= 227 * 1024, "score kernel: shared memory budget");
static_assert(sI0 - sE0 >= 3 * 4096 * 4, "the three d(blend) planes fit between e0 and the arg-max bytes");
static_assert(sE0 % 16 == 0 && sDY3 % 16 == 0 && sA8 % 16 == 0 && sX3 % 16 == 0 && sM2 % 16 == 0 && sW % 16 == 0, "alignment");

struct ScoreParams {
  const uint8_t *framesA, *framesB;
  const float* z;
  const uint2* pack;
  const float *b0, *b1, *b2, *b3, *w4, *b4, *wl1, *bl1, *wl2, *bl2;
  const float* target_replace;       // NULL: pass 0 computes negpred = critic(B) in this kernel
  const float* target_inject;        // NULL: no inject pass
  const float* masks[3][3];          // forced dropout masks [pass 0/1/2][m_e2, m_e3, m_v]; all NULL: none / drawn
  unsigned long long seed;
  unsigned long long* rng_state;
  float p_drop, keep;
  const int* roll_dev;
  int B, roll;
  const float* vpred;
  float l1, l2, reg_scale, gscale, inv_n;
  float *negpred, *pred_replace, *pred_inject, *losses, *dz;
};

// correctly rounded b / 255.0f for b = 0..255 (checked exhaustively): one Newton step on b * fl(1/255)
__device__ __forceinline__ float div255(float b) {
  constexpr float k = 1.f / 255.f;
  const float q = __fmul_rn(b, k);
  return fmaf(fmaf(-255.f, q, b), k, q);
}

__device__ long long* g_hgs_trace = nullptr;
// debug: clock64() at the phase boundaries of the two scoring passes of CTA 0's first two frames (tools/hg_trace.py)
#define HS_MARK(k)                                                                                       \
  do {                                                                                                   \
    if (trace && tid == 0 && fr < 2 && pass >= 1) trace[fr * 32 + 16 * (pass - 1) + (k)] = clock64();     \
  } while (0)

__global__ void __launch_bounds__(NT, 1) hg_score_kernel(const ScoreParams p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, odd = g & 1;
  const int lj = lane >> 3, lr = lane & 7, pixoff = lr + 8 * (lj & 1), chunk = lj >> 1;
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(smraw);
  const uint2* sWf = reinterpret_cast<const uint2*>(smraw + sW);
  float* sBi = reinterpret_cast<float*>(smraw + sBias);
  float* sHWf = reinterpret_cast<float*>(smraw + sHW);
  float* sH = reinterpret_cast<float*>(smraw + sVec);
  float *sV = sH + 32, *sDH = sH + 64, *sDV = sH + 96, *sNeg = sH + 128;
  float* fX3 = reinterpret_cast<float*>(smraw + sX3);
  float* fM2 = reinterpret_cast<float*>(smraw + sM2);
  float* fM3 = reinterpret_cast<float*>(smraw + sM3);
  float* fMV = reinterpret_cast<float*>(smraw + sMV);
  uint8_t *bI0 = smraw + sI0, *bI1 = smraw + sI1, *bI2 = smraw + sI2, *bI3 = smraw + sI3;
  const uint8_t *bA8 = smraw + sA8, *bB8 = smraw + sB8;
  const float* fZ = reinterpret_cast<const float*>(smraw + sZ);
  float* fDZ = reinterpret_cast<float*>(smraw + sDZ);
  float* fDX = reinterpret_cast<float*>(smraw + sE0);           // d(blend), three fp32 planes, during the last phase of a pass
  long long* trace = blockIdx.x == 0 ? g_hgs_trace : nullptr;
  int fr = 0;
  __nv_bfloat16* hE0 = reinterpret_cast<__nv_bfloat16*>(smraw + sE0);
  __nv_bfloat16* hE1 = reinterpret_cast<__nv_bfloat16*>(smraw + sE1);
  __nv_bfloat16* hE2 = reinterpret_cast<__nv_bfloat16*>(smraw + sE2);
  __nv_bfloat16* hDY0 = reinterpret_cast<__nv_bfloat16*>(smraw + sX);
  __nv_bfloat16* hDY1 = reinterpret_cast<__nv_bfloat16*>(smraw + sDY1);
  __nv_bfloat16* hDY2 = reinterpret_cast<__nv_bfloat16*>(smraw + sDY2);
  __nv_bfloat16* hDY3 = reinterpret_cast<__nv_bfloat16*>(smraw + sDY3);

  const unsigned long long rng_call = p.rng_state ? p.rng_state[0] : 0ull;
  const bool forced = p.masks[1][0] != nullptr;
  // ---- prologue: zero every plane (halos stay zero), weight fragments, biases, head weights
  {
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int e = tid; e < sI0 / 16; e += NT) reinterpret_cast<uint4*>(smraw)[e] = z4;
    const uint4* src = reinterpret_cast<const uint4*>(p.pack);
    for (int e = tid; e < F_D2 * 16; e += NT) reinterpret_cast<uint4*>(smraw + sW)[e] = __ldg(src + e);
    for (int e = tid; e < (NSTEPS - B_C3D) * 16; e += NT) reinterpret_cast<uint4*>(smraw + sW + F_D2 * 256)[e] = __ldg(src + B_C3D * 16 + e);
  }
  if (tid < 8) { sBi[tid] = __ldg(p.b0 + tid); sBi[8 + tid] = __ldg(p.b1 + tid); sBi[16 + tid] = __ldg(p.b2 + tid); }
  if (tid < 16) sBi[24 + tid] = __ldg(p.b3 + tid);
  for (int e = tid; e < 1024; e += NT) sHWf[hWl1 + e] = __ldg(p.wl1 + e);
  if (tid < 32) { sHWf[hBl1 + tid] = __ldg(p.bl1 + tid); sHWf[hWl2 + tid] = __ldg(p.wl2 + tid); sHWf[hB4 + tid] = __ldg(p.b4 + tid); }
  if (tid == 0) sHWf[hBl2] = __ldg(p.bl2);
  if (!forced && !p.rng_state) {                       // no dropout: identity masks, written once
    fM2[tid] = 1.f;
    if (tid < 256) fM3[tid] = 1.f;
    if (tid < 32) fMV[tid] = 1.f;
  }
  int roll = p.roll_dev ? *p.roll_dev : p.roll;
  roll = ((roll % 64) + 64) & 63;
  const int pass_lo = p.target_replace ? 1 : 0, pass_hi = p.target_inject ? 3 : 2;
  float loss_r = 0.f, loss_i = 0.f, reg1 = 0.f, reg2 = 0.f;

  for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
    __syncthreads();                                   // the previous frame's last pass is done with the raw bytes
    {
      const uint8_t *srcA = p.framesA + (size_t)n * 12288, *srcB = p.framesB + (size_t)n * 12288;
      const float* srcZ = p.z + (size_t)n * 4096;
      for (int c = tid; c < 768; c += NT) { cp_async16(smb + sA8 + c * 16, srcA + c * 16); cp_async16(smb + sB8 + c * 16, srcB + c * 16); }
      for (int c = tid; c < 1024; c += NT) cp_async16(smb + sZ + c * 16, srcZ + c * 4);
      cp_async_commit();
    }
    for (int pass = pass_lo; pass < pass_hi; ++pass) {
      // ================= this pass's dropout masks; then the frame / blend -> pair-duplicated bf16 plane
      if (forced) {
        const float *q2 = p.masks[pass][0] + (size_t)n * 512, *q3 = p.masks[pass][1] + (size_t)n * 256, *qv = p.masks[pass][2] + (size_t)n * 32;
        if (tid < 128) reinterpret_cast<float4*>(fM2)[tid] = __ldg(reinterpret_cast<const float4*>(q2) + tid);
        else if (tid < 192) reinterpret_cast<float4*>(fM3)[tid - 128] = __ldg(reinterpret_cast<const float4*>(q3) + tid - 128);
        else if (tid < 200) reinterpret_cast<float4*>(fMV)[tid - 192] = __ldg(reinterpret_cast<const float4*>(qv) + tid - 192);
      } else if (p.rng_state) {
        draw_masks3(p.seed, rng_call + (unsigned long long)(pass - pass_lo), p.B, n, p.p_drop, p.keep, fM2, fM3, fMV, tid);
      }
      cp_async_wait_all();
      __syncthreads();
      HS_MARK(0);
      {
        // e0 / e1 / e2: the previous pass parked d(blend) over them, halos included: clear them (interiors are rewritten below)
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int e = tid; e < PB1 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sE0)[e] = z4;
        for (int e = tid; e < PB2 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sE1)[e] = z4;
        for (int e = tid; e < PB3 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sE2)[e] = z4;
      }
      for (int e = tid; e < 4096; e += NT) {
        const int y = e >> 6, x = e & 63;
        const uint8_t* a = bA8 + (y * 64 + ((x + roll) & 63)) * 3;    // shift_batch rolls A only (main.py:355-357)
        const uint8_t* b = bB8 + e * 3;
        float v[3];
        if (pass == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) v[c] = div255((float)b[c]);
        } else {
          // formed in fp32 exactly as the reference does: u8 -> float / 255, two products, one sum, no contraction
          const float zz = fZ[e], omz = __fsub_rn(1.f, zz);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float fa = div255((float)a[c]), fb = div255((float)b[c]);
            const float keepv = pass == 1 ? fa : fb, put = pass == 1 ? fb : fa;
            v[c] = __fadd_rn(__fmul_rn(keepv, omz), __fmul_rn(zz, put));
          }
        }
        const uint2 q = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], 0.f));
        uint8_t* row = smraw + sX + (size_t)(y + 1) * (PX * 16);
        *reinterpret_cast<uint2*>(row + (x + 1) * 16) = q;
        *reinterpret_cast<uint2*>(row + x * 16 + 8) = q;
        // the high half of entry 64 is halo (pixel 65): this region held the previous pass's gradient plane
        if (x == 63) *reinterpret_cast<uint2*>(row + 64 * 16 + 8) = make_uint2(0u, 0u);
      }
      __syncthreads();
      const float ytgt = pass == 0 ? 0.f : (pass == 1 ? (p.target_replace ? __ldg(p.target_replace + n) : sNeg[0]) : __ldg(p.target_inject + n));

      HS_MARK(1);
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
      HS_MARK(2);
      // ================= F1: features.3 (8 -> 8) on 32x32 -> e1; the scatter targets of the backward are cleared meanwhile
      {
        if (pass != 0) {
          const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
          for (int e = tid; e < PBX / 16; e += NT) reinterpret_cast<uint4*>(smraw + sX)[e] = z4;
          for (int e = tid; e < PB1 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sDY1)[e] = z4;
          for (int e = tid; e < PB2 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sDY2)[e] = z4;
          for (int e = tid; e < 2 * PB3 / 16; e += NT) reinterpret_cast<uint4*>(smraw + sDY3)[e] = z4;
        }
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
      HS_MARK(3);
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
      HS_MARK(4);
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
      HS_MARK(5);
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
      // ================= F6: Dropout, crit.4 Linear(32,1), Sigmoid, MSE and its gradient
      float w4c[16];
      if (pass != 0) {
        const float* wc = p.w4 + ((tid & 1) * 16) * 256 + (tid >> 1);
#pragma unroll
        for (int i = 0; i < 16; ++i) w4c[i] = __ldg(wc + i * 256);
      }
      if (warp == 0) {
        const float wk = sHWf[hWl2 + lane], vm = sV[lane] * fMV[lane];
        const float zz = warp_sum(wk * vm) + sHWf[hBl2];
        const float pr = sigmoidf_(zz);
        if (pass == 0) {
          if (lane == 0) { sNeg[0] = pr; p.negpred[n] = pr; }
        } else {
          const float d = pr - ytgt;
          if (pass == 1) loss_r = fmaf(d, d, loss_r); else loss_i = fmaf(d, d, loss_i);
          if (lane == 0) (pass == 1 ? p.pred_replace : p.pred_inject)[n] = pr;
          const float dl = p.gscale * 2.f * d * pr * (1.f - pr);
          sDV[lane] = sV[lane] > 0.f ? dl * wk * fMV[lane] : 0.f;
        }
      }
      __syncthreads();
      if (pass == 0) continue;                           // forward only: negpred
      HS_MARK(6);
      // ================= B5: crit.1 backward
      {
        const int k = tid >> 4, part = tid & 15;
        float s = sHWf[hWl1 + (2 * part) * 32 + k] * sDV[2 * part] + sHWf[hWl1 + (2 * part + 1) * 32 + k] * sDV[2 * part + 1];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (part == 0) sDH[k] = sH[k] > 0.f ? s : 0.f;
      }
      __syncthreads();
      // ================= B4: features.14 backward, Dropout + pool + ReLU backward -> d(features.10 output)
      {
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
      HS_MARK(7);
      // ================= B3: features.10 input gradient (16 -> 8 on 8x8), Dropout + pool + ReLU backward -> d(features.6 output)
      if (warp < 4) {
        const int mt = warp;
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
      HS_MARK(8);
      // ================= B2: features.6 input gradient (16x16), pool + ReLU backward -> d(features.3 output)
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
      }
      __syncthreads();
      HS_MARK(9);
      // ================= B1: features.3 input gradient (32x32), pool + ReLU backward -> d(features.0 output) (64x64x8, region X)
      {
        const int x0 = (warp & 1) * 16, r0 = (warp >> 1) * 4;
        uint2 w[3][2][1];
#pragma unroll
        for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(W_BWD + B_C1D - B_C3D + s) * 32 + lane];
        const uint32_t aA = smb + sDY1 + (uint32_t)((r0 * P1 + x0 + pixoff + chunk) * 16);
        const uint32_t aB = smb + sDY1 + (uint32_t)((r0 * P1 + x0 + pixoff + 2) * 16);
        slide_bf<4, 2, 1>(
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
                for (int q = 0; q < 4; ++q) {
                  const int y = r0 + e + r, x = x0 + g + 8 * (q >> 1), ci = 2 * t + (q & 1), idx = bI0[(y * 32 + x) * 8 + ci];
                  if (idx < 4)
                    hDY0[((2 * y + (idx >> 1) + 1) * PX + 2 * x + (idx & 1) + 1) * 8 + ci] = __float2bfloat16_rn(r ? bot[q] : top[q]);
                }
            });
      }
      __syncthreads();
      HS_MARK(10);
      // ================= B0: features.0 input gradient (8 -> 3 on 64x64), contracted with (B - A) / (A - B) on the way out:
      // d loss / d Z = sum_c dX[c] * (put - keep)[c]  (main.py:395, 406); the regulariser's gradient rides on pass 1
      {
        const int x0 = (warp & 3) * 16, r0 = (warp >> 2) * 16;
        uint2 w[3][2][1];
#pragma unroll
        for (int s = 0; s < 6; ++s) w[s >> 1][s & 1][0] = sWf[(W_BWD + B_C0D - B_C3D + s) * 32 + lane];
        const uint32_t aA = smb + sX + (uint32_t)((r0 * PX + x0 + pixoff + chunk) * 16);
        const uint32_t aB = smb + sX + (uint32_t)((r0 * PX + x0 + pixoff + 2) * 16);
        slide_bf<16, 2, 1>(
            w,
            [&](int i, uint32_t(&a)[2][4]) {
              ldsm4(a[0], aA + (uint32_t)(i * (PX * 16)));
              ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (PX * 16)));
              a[1][2] = a[1][3] = 0u;
            },
            [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
              // columns 0..2 of the 8-wide tile are the three input channels: lanes t = 0 (channels 0, 1) and t = 1 (channel 2)
              // park them as three fp32 planes (e0 .. d(features.10 output) are dead by now)
              if (t < 2) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const int pix = (r0 + e + r) * 64 + x0 + g + 8 * h;
                    fDX[(2 * t) * 4096 + pix] = r ? bot[2 * h] : top[2 * h];
                    if (t == 0) fDX[4096 + pix] = r ? bot[2 * h + 1] : top[2 * h + 1];
                  }
              }
            });
      }
      __syncthreads();
      // d loss / d Z = sum_c dX[c] * (put - keep)[c] (main.py:395, 406), every thread a pixel; the regulariser's gradient
      // (main.py:415-429) rides on the first scoring pass
      {
        const float vf = p.vpred ? 1.f - __ldg(p.vpred + n) : 1.f;
        const float sgn = pass == 2 ? -(1.f / 255.f) : (1.f / 255.f);
        for (int e = tid; e < 4096; e += NT) {
          const int y = e >> 6, x = e & 63;
          const uint8_t* pa = bA8 + (y * 64 + ((x + roll) & 63)) * 3;
          const uint8_t* pb = bB8 + e * 3;
          float s = fDX[e] * (float)((int)pb[0] - (int)pa[0]);
          s = fmaf(fDX[4096 + e], (float)((int)pb[1] - (int)pa[1]), s);
          s = fmaf(fDX[8192 + e], (float)((int)pb[2] - (int)pa[2]), s);
          s *= sgn;
          if (pass == 1) {
            const float u = vf * fZ[e];
            reg1 += fabsf(u); reg2 = fmaf(u, u, reg2);
            const float sg = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
            fDZ[e] = s + p.reg_scale * vf * (p.l1 * sg + 2.f * p.l2 * u);
          } else {
            fDZ[e] += s;
          }
        }
      }
      __syncthreads();
      HS_MARK(11);
    }
    // the frame's d loss / d mask: one coalesced pass
    {
      float4* d = reinterpret_cast<float4*>(p.dz + (size_t)n * 4096);
      const float4* sdz = reinterpret_cast<const float4*>(fDZ);
      for (int e = tid; e < 1024; e += NT) d[e] = sdz[e];
    }
    ++fr;
  }
  cp_async_wait_all();
  // ---- loss terms: one atomic per warp / CTA into the four scalars the host zeroed
  if (warp == 0 && lane == 0) {
    atomicAdd(p.losses, loss_r * p.inv_n);
    if (p.target_inject) atomicAdd(p.losses + 1, loss_i * p.inv_n);
  }
  reg1 = warp_sum(reg1); reg2 = warp_sum(reg2);
  if (lane == 0 && p.l1 != 0.f) atomicAdd(p.losses + 2, reg1 * p.l1 * p.inv_n * (1.f / 4096.f));
  if (lane == 0 && p.l2 != 0.f) atomicAdd(p.losses + 3, reg2 * p.l2 * p.inv_n * (1.f / 4096.f));
  if (p.rng_state && tid == 0) {                     // last CTA to finish advances the call counter (every CTA has read it)
    __threadfence();
    if (atomicAdd(&p.rng_state[1], 1ull) == gridDim.x - 1) {
      p.rng_state[1] = 0;
      p.rng_state[0] = rng_call + (unsigned long long)(pass_hi - pass_lo);
    }
  }
}

}  // namespace hs
}  // namespace cgs

using namespace cgs;

namespace cgs { namespace hs { int set_trace(long long* b) { return cudaMemcpyToSymbol(g_hgs_trace, &b, sizeof(b)) == cudaSuccess ? 0 : -2; } } }

extern "C" int cgs_hg_score_bf16(const uint8_t* frames_a, const uint8_t* frames_b, int32_t B, int32_t roll, const int32_t* roll_dev,
                                 const float* z, const float* target_replace, const float* target_inject, const float* const* masks9,
                                 float p_drop, uint64_t seed, uint64_t* rng_state, const cgs_critic_weights* w, const uint32_t* pack,
                                 float loss_grad, const float* vpred, float l1, float l2, float* negpred, float* pred_replace,
                                 float* pred_inject, float* losses, float* dz, void* stream) {
  CGS_REQUIRE(frames_a && frames_b && z && w && pack && pred_replace && losses && dz && B > 0, "hg_score_bf16: bad args");
  CGS_REQUIRE(target_replace || negpred, "hg_score_bf16: without target_replace the kernel computes negpred and needs its output buffer");
  CGS_REQUIRE(!target_inject || pred_inject, "hg_score_bf16: the inject pass needs pred_inject");
  CGS_REQUIRE((((uintptr_t)frames_a | (uintptr_t)frames_b | (uintptr_t)pack) & 15) == 0, "hg_score_bf16: frames and pack must be 16-byte aligned");
  hs::ScoreParams p;
  memset(&p, 0, sizeof(p));
  if (masks9) {
    const int lo = target_replace ? 1 : 0, hi = target_inject ? 3 : 2;
    bool any = false, all = true;
    for (int q = lo; q < hi; ++q)
      for (int k = 0; k < 3; ++k) {
        const float* m = masks9[q * 3 + k];
        any = any || m; all = all && m;
        CGS_REQUIRE(((uintptr_t)m & 15) == 0, "hg_score_bf16: masks must be 16-byte aligned");
        p.masks[q][k] = m;
      }
    CGS_REQUIRE(!any || all, "hg_score_bf16: forced dropout masks are all-or-none over the passes that run");
    if (!all) memset(p.masks, 0, sizeof(p.masks));
    else if (lo == 1) for (int k = 0; k < 3; ++k) p.masks[0][k] = nullptr;
  }
  CGS_REQUIRE(!(rng_state && p.masks[1][0]), "hg_score_bf16: pass dropout masks OR an rng state, not both");
  CGS_REQUIRE(!rng_state || (p_drop > 0.f && p_drop < 1.f), "hg_score_bf16: rng dropout needs 0 < p < 1");
  p.framesA = frames_a; p.framesB = frames_b; p.z = z; p.pack = reinterpret_cast<const uint2*>(pack);
  p.b0 = w->b0; p.b1 = w->b1; p.b2 = w->b2; p.b3 = w->b3; p.w4 = w->w4; p.b4 = w->b4;
  p.wl1 = w->wl1; p.bl1 = w->bl1; p.wl2 = w->wl2; p.bl2 = w->bl2;
  CGS_REQUIRE(p.b0 && p.b1 && p.b2 && p.b3 && p.w4 && p.b4 && p.wl1 && p.bl1 && p.wl2 && p.bl2, "hg_score_bf16: NULL parameter tensor");
  p.target_replace = target_replace; p.target_inject = target_inject;
  p.seed = seed; p.rng_state = (unsigned long long*)rng_state; p.p_drop = p_drop; p.keep = rng_state ? 1.f / (1.f - p_drop) : 1.f;
  p.roll_dev = roll_dev; p.B = B; p.roll = roll; p.vpred = vpred; p.l1 = l1; p.l2 = l2;
  p.inv_n = 1.f / (float)B; p.gscale = loss_grad / (float)B; p.reg_scale = loss_grad / ((float)B * 4096.f);
  p.negpred = negpred; p.pred_replace = pred_replace; p.pred_inject = pred_inject; p.losses = losses; p.dz = dz;
  cudaStream_t st = (cudaStream_t)stream;
  cudaFuncSetAttribute(hs::hg_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hs::S_SMEM);
  if (cudaMemsetAsync(losses, 0, 4 * sizeof(float), st) != cudaSuccess) return check_launch("hg_score_bf16.memset");
  const int sms = device_sms(), per = (B + sms - 1) / sms, grid = (B + per - 1) / per;
  hs::hg_score_kernel<<<grid, hg::NT, hs::S_SMEM, st>>>(p);
  return check_launch("hg_score_bf16");
}
