// One critic_pipe training step (reference main.py:185-198) as ONE persistent kernel, chfak = 1:
//   uint8 frame -> /255 + shift_batch roll -> NewCritic forward (nets.py:169-212, dropout masks applied) -> MSE/BCE loss
//   -> full backward -> parameter gradients accumulated into the caller's (flat Adam) gradient tensors.
//
// Why: at chfak 1 the critic is 1.7 MMAC per frame and every activation of a frame fits in shared memory (e0 is 32 KB),
// so the compulsory HBM traffic is the 12 KB uint8 frame and nothing else; the per-layer kernels it replaces moved
// ~0.3 MB per frame through L2/HBM and paid ~10 us of launch/ramp latency per layer at batch 256.
//
// One CTA (16 warps) owns a frame at a time and walks the layers with every operand in shared memory:
//   * convolutions (fprop and dgrad) are implicit GEMMs on mma.sync.m16n8k8 TF32 (fp32 accumulate): an M-tile is 16
//     pixels of one row, K = the 8 input channels of one filter tap, N = 8 output channels.  Activations live in smem
//     as two 4-channel half-planes so that ONE ldmatrix.x4 fetches a whole A fragment conflict-free, and a warp slides
//     down a 16-pixel column strip so every fragment is reused by the three filter rows (3 ldmatrix per 9 MMAs);
//   * ReLU + 2x2 max-pool + first-max arg-max happen on the accumulator fragments (row pairs live in one warp, the
//     x-neighbour is one shuffle away); the backward scatter (pool/ReLU backward) is the dgrad epilogue;
//   * weight gradients are GEMMs with K = pixels: A = gathered input taps, B = output gradient, accumulated per CTA
//     in shared memory / registers over all its frames and pushed with ONE round of REDs per CTA;
//   * the 4x4 bottleneck conv and the MLP head are FFMA on the tiny vectors.
// tcgen05 is not used here on purpose: every GEMM has N = 8..16 and M-tiles of 16 pixels; a 128-row UMMA tile would
// be >85 % padding and its operands would have to be re-laid out in the canonical layout per tap.
#include <string.h>
#include "critic_tail.cuh"

namespace cgs {
namespace cf {

constexpr int NT = 512;
// haloed planes: pitch in pixels, half-plane size in floats (4 channels per pixel).  PL1, PL2 == 4 (mod 32) banks (the
// weight-gradient gathers walk pixels 2t.., 8 banks apart, with the two half-planes 4 banks apart); PL3 == 16 (mod 32)
constexpr int P0 = 66, SX = 66 * 66 * 4;          // RGB0 frame, one 16-byte pixel
constexpr int P1 = 34, PL1 = 34 * 34 * 4 + 20;    // 32x32 maps
constexpr int P2 = 18, PL2 = 18 * 18 * 4 + 20;    // 16x16 maps
constexpr int P3 = 10, PL3 = 10 * 10 * 4;         // 8x8 maps
static_assert(PL1 % 32 == 4 && PL2 % 32 == 4 && PL3 % 32 == 16, "bank layout of the half-planes");
// ---- shared memory map (float offsets)
constexpr int oA = 0;                              // region A: e0 | dY1      (B0: the re-staged frame)
constexpr int oE0 = oA, oDY1 = oA + 2 * PL1, oXB = oA;
constexpr int szA = 4 * PL1;                       // 18496 >= SX
constexpr int oB = oA + szA;                       // region B: the frame (F0) | dE0, dY2, dY3 (backward)
constexpr int oX = oB, oDE0 = oB, oDY2 = oDE0 + 8192, oDY3 = oDY2 + 2 * PL2;
constexpr int szB = SX;
constexpr int oI0 = oB + szB;                      // bytes [32*32][8]
constexpr int oE1 = oI0 + 2048;
constexpr int oI1 = oE1 + 2 * PL2;                 // bytes [16*16][8]
constexpr int oE2 = oI1 + 512;                     // e2 * dropout mask (the operand of features.10)
constexpr int oI2 = oE2 + 2 * PL3;                 // bytes [8*8][8]
constexpr int oM2 = oI2 + 128;
constexpr int oX3 = oM2 + 512;                     // e3 * mask in the 4x4 conv's K order (ci*16 + pixel)
constexpr int oI3 = oX3 + 256;                     // bytes [4*4][16]
constexpr int oM3 = oI3 + 64;
constexpr int oHead = oM3 + 256;                   // h[32] v[32] mv[32] dh[32] dv[32]
constexpr int oU8 = oHead + 256;                   // raw frame bytes (12288)
constexpr int oAcc = oU8 + 3072;
constexpr int oW = oAcc + szAcc;                   // weight fragments, [step][lane][2]
constexpr int wL0 = 0, wL1f = 384, wL2f = wL1f + 576, wL3f = wL2f + 576, wL3d = wL3f + 1152, wL2d = wL3d + 1152,
              wL1d = wL2d + 576, szW = wL1d + 576;
constexpr int oBias = oW + szW;                    // b0[8] b1[8] b2[8] b3[16]
constexpr int oHW = oBias + 64;                    // head weights: wl1[1024] bl1[32] wl2[32] bl2[1] b4[32]
constexpr int hWl1 = 0, hBl1 = 1024, hWl2 = 1056, hBl2 = 1088, hB4 = 1092, szHW = 1124;
constexpr int SMEM_FLOATS = oHW + szHW;
static_assert(szA >= SX, "region A must hold the re-staged frame");
static_assert(oDY3 + 4 * PL3 <= oB + szB, "region B overflow");
static_assert(SMEM_FLOATS * 4 <= 227 * 1024, "shared memory budget");


__device__ long long* g_cf_trace = nullptr;   // debug: clock64() at every phase boundary of CTA 0 (tools/fused_trace.py)
#define CF_MARK(k)                                                       \
  do {                                                                   \
    if (trace && tid == 0) trace[fr * 24 + (k)] = clock64();             \
  } while (0)

// Weight gradient of an 8-input-channel 3x3 conv over `nks` k-steps of 8 pixels (TW = map width): acc[mt] rows are
// (tap 2mt | tap 2mt+1) x ci; row 8 of mt 4 is the all-ones row (bias gradient).  X: haloed half-planes of the layer input,
// DY: haloed half-planes of the output gradient.  The mma's k index is mapped to pixels as k = t -> 2t, k = t+4 -> 2t+1,
// so a lane needs the 4 consecutive pixels 2t..2t+3 of each filter row for all three kx: 12 loads feed the 9 taps.
template <int TW, int P, int PL>
__device__ __forceinline__ void wgrad8(float (&acc)[5][4], const float* __restrict__ sX, const float* __restrict__ sDY, int ks0,
                                       int nks, int g, int t) {
  const int offX = (g >> 2) * PL + (g & 3) + 8 * t;
  const int offY = offX + (P + 1) * 4;
  const uint32_t ones = g == 0 ? 0x3f800000u : 0u;
  constexpr int KPR = TW / 8;   // k-steps per row
  for (int ks = ks0; ks < ks0 + nks; ++ks) {
    const int y = ks / KPR, x0 = (ks % KPR) * 8;
    const float* px = sX + offX + (y * P + x0) * 4;
    const uint32_t b0 = __float_as_uint(sDY[offY + (y * P + x0) * 4]), b1 = __float_as_uint(sDY[offY + (y * P + x0) * 4 + 4]);
    uint32_t v[3][4];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int i = 0; i < 4; ++i) v[ky][i] = __float_as_uint(px[(ky * P + i) * 4]);
#pragma unroll
    for (int mt = 0; mt < 5; ++mt) {
      constexpr int dummy = 0;
      (void)dummy;
      const int ta = 2 * mt, tb = 2 * mt + 1;
      uint32_t a[4];
      a[0] = v[ta / 3][ta % 3];
      a[2] = v[ta / 3][ta % 3 + 1];
      if (mt < 4) {
        a[1] = v[tb / 3][tb % 3];
        a[3] = v[tb / 3][tb % 3 + 1];
      } else {
        a[1] = a[3] = ones;
      }
      mma_tf32(acc[mt], a, b0, b1);
    }
  }
}

// acc[mt] of wgrad8 -> a per-warp tile laid out like the OIHW gradient + bias: [co*72 + ci*9 + tap | 576 + co]
__device__ __forceinline__ void wgrad8_store(const float (&acc)[5][4], float* tile, int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 5; ++mt)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = 2 * t + (q & 1), tap = 2 * mt + (q >> 1);
      if (tap < 9) tile[(co * 8 + g) * 9 + tap] = acc[mt][q];
      else if (g == 0) tile[576 + co] = acc[mt][q];
    }
}


// raw bytes -> haloed [66][66] x (r,g,b,0) fp32 /255 with the circular W-roll; tf32(b * (1/255)) == tf32(b / 255) for all b
__device__ __forceinline__ void stage_frame(const uint8_t* __restrict__ sU8, float* __restrict__ sXd, int roll, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = tid + NT * i, y = p >> 6, x = p & 63;
    const uint8_t* s = sU8 + (y * 64 + ((x + roll) & 63)) * 3;
    float4 v;
    v.x = u8_to_tf32_unit(s[0]); v.y = u8_to_tf32_unit(s[1]); v.z = u8_to_tf32_unit(s[2]); v.w = 0.f;
    *reinterpret_cast<float4*>(sXd + ((y + 1) * P0 + x + 1) * 4) = v;
  }
  if (tid < 260) {   // halo ring
    int y, x;
    if (tid < 66) { y = 0; x = tid; }
    else if (tid < 132) { y = 65; x = tid - 66; }
    else if (tid < 196) { y = tid - 131; x = 0; }
    else { y = tid - 195; x = 65; }
    *reinterpret_cast<float4*>(sXd + (y * P0 + x) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// The three dropout masks of frame n, drawn exactly as cgs_dropout_masks would fill [B*512 | B*256 | B*32] floats:
// 200 threads, one Philox call (4 Bernoulli draws) each.
__device__ __forceinline__ void draw_masks(const Params& p, int n, unsigned long long call, float* sm, int tid) {
  if (tid >= 200) return;
  long long vec;
  float* dst;
  if (tid < 128) { vec = (long long)n * 128 + tid; dst = sm + oM2 + tid * 4; }
  else if (tid < 192) { vec = (long long)p.B * 128 + (long long)n * 64 + (tid - 128); dst = sm + oM3 + (tid - 128) * 4; }
  else { vec = (long long)p.B * 192 + (long long)n * 8 + (tid - 192); dst = sm + oHead + 64 + (tid - 192) * 4; }
  uint32_t c[4] = {(uint32_t)vec, (uint32_t)(vec >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
  philox4x32_10(c, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  float4 v;
  v.x = ((float)(c[0] >> 8) * (1.0f / 16777216.0f) >= p.p_drop) ? p.keep : 0.f;
  v.y = ((float)(c[1] >> 8) * (1.0f / 16777216.0f) >= p.p_drop) ? p.keep : 0.f;
  v.z = ((float)(c[2] >> 8) * (1.0f / 16777216.0f) >= p.p_drop) ? p.keep : 0.f;
  v.w = ((float)(c[3] >> 8) * (1.0f / 16777216.0f) >= p.p_drop) ? p.keep : 0.f;
  *reinterpret_cast<float4*>(dst) = v;
}

// cp.async: the raw frame (12288 B) and its dropout masks (512 + 256 + 32 floats) for frame n
__device__ __forceinline__ void prefetch_frame(const Params& p, int n, float* sm, int tid) {
  if (p.frames) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(sm + oU8);
    const uint8_t* src = p.frames + (size_t)n * 12288;
    for (int c = tid; c < 768; c += NT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + c * 16), "l"(src + c * 16));
  }
  if (p.m2) {
    const uint32_t dm = (uint32_t)__cvta_generic_to_shared(sm);
    if (tid < 128) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + (oM2 + tid * 4) * 4), "l"(p.m2 + (size_t)n * 512 + tid * 4));
    else if (tid < 192) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + (oM3 + (tid - 128) * 4) * 4), "l"(p.m3 + (size_t)n * 256 + (tid - 128) * 4));
    else if (tid < 200) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + (oHead + 64 + (tid - 192) * 4) * 4), "l"(p.mv + (size_t)n * 32 + (tid - 192) * 4));
  }
  asm volatile("cp.async.commit_group;\n" ::);
}

// fp32 NHWC frame [64][64][3] in global memory -> haloed [66][66] x (r,g,b,0), TF32-rounded (input-gradient mode)
__device__ __forceinline__ void stage_frame_f32(const float* __restrict__ src, float* __restrict__ sXd, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = tid + NT * i, y = p >> 6, x = p & 63;
    const float* s = src + p * 3;
    *reinterpret_cast<float4*>(sXd + ((y + 1) * P0 + x + 1) * 4) = make_float4(tf32r(__ldg(s)), tf32r(__ldg(s + 1)), tf32r(__ldg(s + 2)), 0.f);
  }
  if (tid < 260) {
    int y, x;
    if (tid < 66) { y = 0; x = tid; }
    else if (tid < 132) { y = 65; x = tid - 66; }
    else if (tid < 196) { y = tid - 131; x = 0; }
    else { y = tid - 195; x = 65; }
    *reinterpret_cast<float4*>(sXd + (y * P0 + x) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// MODE 3 staging: the occlusion blend of main.py:395 (which = 0: A*(1-Z) + Z*B) or :406 (which = 1: B*(1-Z) + Z*A), formed in
// fp32 exactly as the reference does (u8 -> float / 255, two products, one sum, no contraction), then TF32-rounded into the
// haloed (r,g,b,0) tile.  sA8 is read with the shift_batch roll, sB8 without (main.py:355-357).
__device__ __forceinline__ void stage_blend(const uint8_t* __restrict__ sA8, const uint8_t* __restrict__ sB8, const float* __restrict__ z,
                                            int which, int roll, float* __restrict__ sXd, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = tid + NT * i, y = p >> 6, x = p & 63;
    const uint8_t* a = sA8 + (y * 64 + ((x + roll) & 63)) * 3;
    const uint8_t* b = sB8 + p * 3;
    const float zz = __ldg(z + p), omz = __fsub_rn(1.f, zz);
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float fa = __fdiv_rn((float)a[c], 255.f), fb = __fdiv_rn((float)b[c], 255.f);
      const float keep = which ? fb : fa, put = which ? fa : fb;
      v[c] = tf32r(__fadd_rn(__fmul_rn(keep, omz), __fmul_rn(zz, put)));
    }
    *reinterpret_cast<float4*>(sXd + ((y + 1) * P0 + x + 1) * 4) = make_float4(v[0], v[1], v[2], 0.f);
  }
  if (tid < 260) {
    int y, x;
    if (tid < 66) { y = 0; x = tid; }
    else if (tid < 132) { y = 65; x = tid - 66; }
    else if (tid < 196) { y = tid - 131; x = 0; }
    else { y = tid - 195; x = 65; }
    *reinterpret_cast<float4*>(sXd + (y * P0 + x) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// XG = false: one critic_pipe training step (weight gradients, optimizer).  XG = true: the same forward and loss with FROZEN
// parameters and the gradient w.r.t. the input frame instead (critic(replaced) / critic(injected) of the Hourglass loop,
// main.py:396-411, and the saliency baseline, main.py:949-951): all weight-gradient work is compiled out and the backward ends
// with features.0's input gradient, computed in two 32-row bands from the arg-max-tagged pooled gradient.
// MODE 2 = MODE 1's forward only (pred for fp32 frames, e.g. `negpred = critic(B)` under no_grad, main.py:365-367).
template <int MODE>
__global__ void __launch_bounds__(NT, 1) critic_fused_kernel(const Params p) {
  constexpr bool XG = MODE != 0, FWD = MODE == 2, BL = MODE == 3;
  extern __shared__ __align__(128) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int lj = lane >> 3, lr = lane & 7;               // ldmatrix: this lane addresses row lr of matrix lj
  uint8_t* sI0 = reinterpret_cast<uint8_t*>(sm + oI0);
  uint8_t* sI1 = reinterpret_cast<uint8_t*>(sm + oI1);
  uint8_t* sI2 = reinterpret_cast<uint8_t*>(sm + oI2);
  uint8_t* sI3 = reinterpret_cast<uint8_t*>(sm + oI3);
  float *sH = sm + oHead, *sV = sH + 32, *sMV = sH + 64, *sDH = sH + 96, *sDV = sH + 128;
  float* sAcc = sm + oAcc;
  const float2* sWf = reinterpret_cast<const float2*>(sm + oW);
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(sm);

  long long* trace = blockIdx.x == 0 ? g_cf_trace : nullptr;
  // debug: every CTA's start / end-of-frames / end time in ns (globaltimer) + its SM id, after the 4*24 phase marks
  long long* ctat = g_cf_trace ? g_cf_trace + 96 + blockIdx.x * 4 : nullptr;
  if (ctat && threadIdx.x == 0) {
    long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    ctat[0] = gt; ctat[3] = smid;
  }
  int fr = 0;
  CF_MARK(23);
  const unsigned long long rng_call = p.rng_state ? p.rng_state[0] : 0ull;
  unsigned bar_gen = 0;
  if (p.adam_p && tid == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(bar_gen) : "l"(p.bar + 1) : "memory");
  if (!BL && blockIdx.x < p.B) {
    prefetch_frame(p, blockIdx.x, sm, tid);
    if (p.rng_state) draw_masks(p, blockIdx.x, rng_call, sm, tid);
  }

  // ---- prologue: accumulators, halos that stay zero, weight fragments (TF32, in mma B-fragment order), biases
  if (!XG) {
    for (int e = tid; e < szAcc; e += NT) sAcc[e] = 0.f;
  } else {
    // the accumulator region is free in this mode: it holds features.0's dgrad fragments, B[k = co][n = ci] = W0[co][ci][8 - tap']
    for (int e = tid; e < 9 * 32; e += NT) {
      const int ln = e & 31, gg = ln >> 2, tt = ln & 3, tap = 8 - (e >> 5);
      sAcc[2 * e] = gg < 3 ? tf32r(__ldg(p.w0 + tt * 27 + gg * 9 + tap)) : 0.f;
      sAcc[2 * e + 1] = gg < 3 ? tf32r(__ldg(p.w0 + (tt + 4) * 27 + gg * 9 + tap)) : 0.f;
    }
  }
  for (int e = tid; e < 2 * PL2 + 512 + 2 * PL3; e += NT) sm[oE1 + e] = 0.f;        // e1, idx1, e2 (halos stay zero)
  {
    // 2496 fragment entries, 5 per thread: first all source addresses, then all loads in flight together (a cold launch
    // pays ONE HBM round trip for its weights, not five), then round + store
    const float* src[5][2];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int e = tid + NT * it;
      const int ln = e & 31, gg = ln >> 2, tt = ln & 3;
      const float *x = nullptr, *y = nullptr;
      int s = e >> 5;
      if (e >= szW / 2) {
      } else if (s < 6) {                          // L0 fprop: step (ky, kk): kk 0 = taps (ky,0 | ky,1), kk 1 = tap (ky,2) | 0
        const int ky = s >> 1, kk = s & 1;
        if (tt < 3) {
          x = p.w0 + gg * 27 + tt * 9 + ky * 3 + (kk ? 2 : 0);
          if (!kk) y = p.w0 + gg * 27 + tt * 9 + ky * 3 + 1;
        }
      } else if ((s -= 6) < 18) {                  // L1 / L2 fprop: tap
        const float* w = s < 9 ? p.w1 : p.w2;
        const int tap = s % 9;
        x = w + (gg * 8 + tt) * 9 + tap; y = w + (gg * 8 + tt + 4) * 9 + tap;
      } else if ((s -= 18) < 18) {                 // L3 fprop: (tap, nt)
        const int tap = s >> 1, co = (s & 1) * 8 + gg;
        x = p.w3 + (co * 8 + tt) * 9 + tap; y = p.w3 + (co * 8 + tt + 4) * 9 + tap;
      } else if ((s -= 18) < 18) {                 // L3 dgrad: (tap', ks): B[k = co][n = ci] = W[co][ci][8 - tap']
        const int tap = 8 - (s >> 1), co = (s & 1) * 8 + tt;
        x = p.w3 + (co * 8 + gg) * 9 + tap; y = p.w3 + ((co + 4) * 8 + gg) * 9 + tap;
      } else {                                     // L2 / L1 dgrad: tap'
        s -= 18;
        const float* w = s < 9 ? p.w2 : p.w1;
        const int tap = 8 - s % 9;
        x = w + (tt * 8 + gg) * 9 + tap; y = w + ((tt + 4) * 8 + gg) * 9 + tap;
      }
      src[it][0] = x; src[it][1] = y;
    }
    float val[5][2];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      val[it][0] = src[it][0] ? __ldg(src[it][0]) : 0.f;
      val[it][1] = src[it][1] ? __ldg(src[it][1]) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int e = tid + NT * it;
      if (e < szW / 2) { sm[oW + 2 * e] = tf32r(val[it][0]); sm[oW + 2 * e + 1] = tf32r(val[it][1]); }
    }
  }
  if (tid < 8) { sm[oBias + tid] = __ldg(p.b0 + tid); sm[oBias + 8 + tid] = __ldg(p.b1 + tid); sm[oBias + 16 + tid] = __ldg(p.b2 + tid); }
  if (tid < 16) sm[oBias + 24 + tid] = __ldg(p.b3 + tid);
  for (int e = tid; e < 1024; e += NT) sm[oHW + hWl1 + e] = __ldg(p.wl1 + e);
  if (tid < 32) {
    sm[oHW + hBl1 + tid] = __ldg(p.bl1 + tid);
    sm[oHW + hWl2 + tid] = __ldg(p.wl2 + tid);
    sm[oHW + hB4 + tid] = __ldg(p.b4 + tid);
  }
  if (tid == 0) sm[oHW + hBl2] = __ldg(p.bl2);
  if (!p.m2 && !p.rng_state) {                     // eval mode / p = 0: identity masks, written once
    sm[oM2 + tid] = 1.f;
    if (tid < 256) sm[oM3 + tid] = 1.f;
    if (tid < 32) sm[oHead + 64 + tid] = 1.f;
  }

  int roll = p.roll_dev ? *p.roll_dev : p.roll;
  roll = ((roll % 64) + 64) & 63;
  // gradient accumulators that live in registers over all frames of this CTA:
  //   accW4: features.14 (each thread owns 16 of the 8192 entries), acc0: features.0 (all warps, K split by rows),
  //   accW: features.3 (warps 0-7) or features.6 (warps 8-15), K split over the 8 warps; bsum: features.0 bias
  float accW4[16], acc0[2][4], accW[5][4], bsum[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 16; ++i) accW4[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc0[i >> 2][i & 3] = 0.f;
#pragma unroll
  for (int i = 0; i < 20; ++i) accW[i >> 2][i & 3] = 0.f;
  float loss_acc = 0.f, loss_acc2 = 0.f, reg1 = 0.f, reg2 = 0.f;
  const int npass = BL && p.target2 ? 2 : 1;
  uint8_t* sA8 = reinterpret_cast<uint8_t*>(sm + oU8);            // MODE 3: raw bytes of the A frame ...
  uint8_t* sB8 = reinterpret_cast<uint8_t*>(sm + oAcc + 576);     // ... and of the B frame (the accumulator region is free)
  const int odd = g & 1;
  // features.0 weight-gradient rows: m = 8G + r; G < 3: filter row ky = G, the (kx, ci) combos except (2, ky);
  // G = 3: the three left-out combos (ky = r, kx = 2, ci = r), rows 3..7 unused.  Within an 8-row group every lane
  // address 4*(t + kx) + ci is a distinct bank or the same word: the gather is conflict-free.
  int offA0[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int G = 2 * mt + h;
      int ky, kx, ci;
      if (G < 3) {
        const int c = g < 6 + G ? g : g + 1;       // skip combo index 6 + ky == (kx 2, ci ky)
        ky = G; kx = c / 3; ci = c - kx * 3;
      } else {
        ky = g < 3 ? g : 0; kx = 2; ci = g < 3 ? g : 0;
      }
      offA0[mt][h] = (ky * P0 + kx) * 4 + ci;
    }
  const int ldoff8 = (lr + 8 * (lj & 1)) * 4;                 // ldmatrix row offset inside an 8-channel half-plane strip

  for (int n = blockIdx.x; n < p.B; n += gridDim.x)
  for (int which = 0; which < npass; ++which) {
    CF_MARK(0);
    // ================= F0a: frame bytes (+ dropout masks) have landed -> fp32 haloed tile; e0 halo
    const float ytgt = __ldg((BL && which ? p.target2 : p.target) + n);
    if (BL) {
      // both raw frames once per frame; this pass's dropout masks (forced: cp.async; else drawn: call index + pass)
      const uint32_t dm = (uint32_t)__cvta_generic_to_shared(sm);
      if (which == 0) {
        const uint8_t *srcA = p.frames + (size_t)n * 12288, *srcB = p.framesB + (size_t)n * 12288;
        for (int c = tid; c < 768; c += NT) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + oU8 * 4 + c * 16), "l"(srcA + c * 16));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + (oAcc + 576) * 4 + c * 16), "l"(srcB + c * 16));
        }
      }
      if (p.m2) {
        const float *q2 = which ? p.m2b : p.m2, *q3 = which ? p.m3b : p.m3, *qv = which ? p.mvb : p.mv;
        if (tid < 128) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + (oM2 + tid * 4) * 4), "l"(q2 + (size_t)n * 512 + tid * 4));
        else if (tid < 192) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + (oM3 + (tid - 128) * 4) * 4), "l"(q3 + (size_t)n * 256 + (tid - 128) * 4));
        else if (tid < 200) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dm + (oHead + 64 + (tid - 192) * 4) * 4), "l"(qv + (size_t)n * 32 + (tid - 192) * 4));
      }
      asm volatile("cp.async.commit_group;\n" ::);
      if (p.rng_state) draw_masks(p, n, rng_call + which, sm, tid);
    }
    asm volatile("cp.async.wait_all;\n" ::);
    __syncthreads();
    if (BL) stage_blend(sA8, sB8, p.zmask + (size_t)n * 4096, which, roll, sm + oX, tid);
    else if (XG && !p.frames) stage_frame_f32(p.xin + (size_t)n * 12288, sm + oX, tid);
    else stage_frame(reinterpret_cast<const uint8_t*>(sm + oU8), sm + oX, roll, tid);
    if (tid < 264) {   // e0 halo ring (region A is reused by the re-staged frame), both half-planes
      const int h = tid >= 132, q = tid - 132 * h;
      int y, x;
      if (q < 34) { y = 0; x = q; }
      else if (q < 68) { y = 33; x = q - 34; }
      else if (q < 100) { y = q - 67; x = 0; }
      else { y = q - 99; x = 33; }
      *reinterpret_cast<float4*>(sm + oE0 + h * PL1 + (y * P1 + x) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    CF_MARK(1);
    // ================= F0: features.0 (3 -> 8) + ReLU + pool : 16 warps = 4 strips x 4 segments of 16 rows
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
      uint8_t* dI = sI0 + ((r0 >> 1) * 32 + ((x0 + g) >> 1)) * 8 + co;
      slide_rows<16, 2>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + i * (P0 * 16));
            ldsm2(a[1][0], a[1][1], aB + i * (P0 * 16));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
            pool2x2(top, bot, bias0, bias1, odd, [&](int h, float v, int idx) {
              dE[((e >> 1) * P1 + 4 * h) * 4] = tf32r(v);
              dI[((e >> 1) * 32 + 4 * h) * 8] = (uint8_t)idx;
            });
          });
    }
    __syncthreads();

    CF_MARK(2);
    // ================= F1: features.3 (8 -> 8) on 32x32 : 2 strips x 8 segments of 4 rows; scatter targets are cleared
    {
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = tid; e < 2 * PL1 / 4; e += NT) reinterpret_cast<float4*>(sm + oDY1)[e] = z4;
      for (int e = tid; e < (2 * PL2 + 4 * PL3) / 4; e += NT) reinterpret_cast<float4*>(sm + oDY2)[e] = z4;
      const int x0 = (warp & 1) * 16, r0 = (warp >> 1) * 4;
      float2 w[3][3];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3] = sWf[(wL1f >> 1) + s * 32 + lane];
      const uint32_t aA = smb + (oE0 + (lj >> 1) * PL1 + (r0 * P1 + x0) * 4 + ldoff8) * 4;
      const float bias0 = sm[oBias + 8 + 2 * t], bias1 = sm[oBias + 8 + 2 * t + 1];
      const int co = 2 * t + odd;
      float* dE = sm + oE1 + (co >> 2) * PL2 + (((r0 >> 1) + 1) * P2 + ((x0 + g) >> 1) + 1) * 4 + (co & 3);
      uint8_t* dI = sI1 + ((r0 >> 1) * 16 + ((x0 + g) >> 1)) * 8 + co;
      slide_rows<4, 3>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (i * P1 + kx) * 16);
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
            pool2x2(top, bot, bias0, bias1, odd, [&](int h, float v, int idx) {
              dE[((e >> 1) * P2 + 4 * h) * 4] = tf32r(v);
              dI[((e >> 1) * 16 + 4 * h) * 8] = (uint8_t)idx;
            });
          });
    }
    __syncthreads();

    CF_MARK(3);
    // ================= F2: features.6 (8 -> 8) on 16x16 + Dropout : 8 warps x 2 rows
    if (warp < 8) {
      const int r0 = warp * 2;
      float2 w[3][3];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3] = sWf[(wL2f >> 1) + s * 32 + lane];
      const uint32_t aA = smb + (oE1 + (lj >> 1) * PL2 + (r0 * P2) * 4 + ldoff8) * 4;
      const float bias0 = sm[oBias + 16 + 2 * t], bias1 = sm[oBias + 16 + 2 * t + 1];
      slide_rows<2, 3>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (i * P2 + kx) * 16);
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
            pool2x2(top, bot, bias0, bias1, odd, [&](int h, float v, int idx) {
              const int py = warp, px = (g >> 1) + 4 * h, co = 2 * t + odd, q = (py * 8 + px) * 8 + co;
              sm[oE2 + (co >> 2) * PL3 + ((py + 1) * P3 + px + 1) * 4 + (co & 3)] = tf32r(v * sm[oM2 + q]);
              sI2[q] = (uint8_t)idx;
            });
          });
    }
    __syncthreads();

    CF_MARK(4);
    // ================= F3: features.10 (8 -> 16) on 8x8 + Dropout : 8 warps = 4 row pairs x 2 channel tiles
    // (all threads first put their slice of the features.14 weights in flight: 32 KB from L2, used in F4)
    // chunk order rotated per thread ((i + part/2) & 3) so that the 128-bit shared reads of x3 in F4 / B4 are conflict-free
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
        const float2 w = sWf[(wL3f >> 1) + (tap * 2 + nt) * 32 + lane];
        mma_tf32(acc, a, __float_as_uint(w.x), __float_as_uint(w.y));
      }
      // rows (2mt, 2mt+1) x pixel g: the x-pair lanes split the two channels
      const int co = nt * 8 + 2 * t + odd;
      const float b0 = sm[oBias + 24 + nt * 8 + 2 * t], b1 = sm[oBias + 24 + nt * 8 + 2 * t + 1];
      const float t0 = acc[0] + b0, t1 = acc[1] + b1, u0 = acc[2] + b0, u1 = acc[3] + b1;
      const float rt = __shfl_xor_sync(0xffffffffu, odd ? t0 : t1, 4), rb = __shfl_xor_sync(0xffffffffu, odd ? u0 : u1, 4);
      const float p0 = odd ? rt : t0, p1 = odd ? t1 : rt, p2 = odd ? rb : u0, p3 = odd ? u1 : rb;
      const float m01 = fmaxf(p0, p1), m23 = fmaxf(p2, p3);
      const int i01 = p1 > p0 ? 1 : 0, i23 = p3 > p2 ? 3 : 2;
      float m = fmaxf(m01, m23);
      int idx = m23 > m01 ? i23 : i01;
      if (!(m > 0.f)) { m = 0.f; idx = 4; }
      const int pp = mt * 4 + (g >> 1);
      sm[oX3 + co * 16 + pp] = m * sm[oM3 + pp * 16 + co];
      sI3[pp * 16 + co] = (uint8_t)idx;
    }
    __syncthreads();

    CF_MARK(5);
    // ================= F4: features.14 (4x4 valid conv = 256 -> 32) + ReLU
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
    CF_MARK(6);
    // ================= F5: crit.1 Linear(32,32) + ReLU
    {
      const int nn = tid >> 4, part = tid & 15;
      const float2 wv = *reinterpret_cast<const float2*>(sm + oHW + hWl1 + nn * 32 + 2 * part);
      float s = wv.x * sH[2 * part] + wv.y * sH[2 * part + 1];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0) sV[nn] = fmaxf(s + sm[oHW + hBl1 + nn], 0.f);
    }
    __syncthreads();
    CF_MARK(7);
    // ================= F6: Dropout, crit.4 Linear(32,1), Sigmoid, loss and its gradient; head weight gradients
    // (meanwhile every thread puts its features.14 column slice in flight for B4)
    float w4c[16];
    {
      const float* wc = p.w4 + ((tid & 1) * 16) * 256 + (tid >> 1);
#pragma unroll
      for (int i = 0; i < 16; ++i) w4c[i] = __ldg(wc + i * 256);
    }
    if (warp == 0) {
      const float wk = sm[oHW + hWl2 + lane], vm = sV[lane] * sMV[lane];
      const float z = warp_sum(wk * vm) + sm[oHW + hBl2];
      const float pr = sigmoidf_(z), y = ytgt;
      float dl;
      if (p.bce == 2) {                 // "loss" = mean(pred): the saliency baseline's pred.mean().backward() (main.py:949)
        loss_acc += pr;
        dl = p.gscale * pr * (1.f - pr);
      } else if (p.bce) {
        loss_acc -= y * fmaxf(logf(pr), -100.f) + (1.f - y) * fmaxf(logf(1.f - pr), -100.f);
        dl = p.gscale * (pr - y) / fmaxf(pr * (1.f - pr), 1e-12f) * pr * (1.f - pr);
      } else {
        if (BL && which) loss_acc2 = fmaf(pr - y, pr - y, loss_acc2);
        else loss_acc = fmaf(pr - y, pr - y, loss_acc);
        dl = p.gscale * 2.f * (pr - y) * pr * (1.f - pr);
      }
      if (lane == 0) (BL && which ? p.pred2 : p.pred)[n] = pr;
      if (!XG) {
        if (lane == 0) sAcc[aBl2] += dl;
        sAcc[aWl2 + lane] += dl * vm;
      }
      sDV[lane] = sV[lane] > 0.f ? dl * wk * sMV[lane] : 0.f;
    }
    __syncthreads();
    if (FWD) {                                       // forward only: the masks are free again, fetch the next frame's
      if (n + (int)gridDim.x < p.B) {
        prefetch_frame(p, n + gridDim.x, sm, tid);
        if (p.rng_state) draw_masks(p, n + gridDim.x, rng_call, sm, tid);
      }
      continue;
    }
    CF_MARK(8);
    // ================= B5: crit.1 backward
    {
      if (!XG) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int e = tid + NT * i;
          sAcc[aWl1 + e] = fmaf(sDV[e >> 5], sH[e & 31], sAcc[aWl1 + e]);
        }
        if (tid < 32) sAcc[aBl1 + tid] += sDV[tid];
      }
      const int k = tid >> 4, part = tid & 15;
      float s = sm[oHW + hWl1 + (2 * part) * 32 + k] * sDV[2 * part] + sm[oHW + hWl1 + (2 * part + 1) * 32 + k] * sDV[2 * part + 1];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0) sDH[k] = sH[k] > 0.f ? s : 0.f;
    }
    __syncthreads();
    CF_MARK(9);
    // ================= B4: features.14 backward (weight gradient in registers), Dropout + pool + ReLU backward -> dY3
    {
      const int nn = tid >> 4, part = tid & 15;
      const float d = sDH[nn];
      if (!XG)
#pragma unroll
      for (int i = 0; i < 4; ++i) {                                         // accW4[4i..] <-> chunk (i + rot4) & 3, as in F4
        const float4 x = *reinterpret_cast<const float4*>(sm + oX3 + part * 16 + ((i + rot4) & 3) * 4);
        accW4[4 * i + 0] = fmaf(d, x.x, accW4[4 * i + 0]); accW4[4 * i + 1] = fmaf(d, x.y, accW4[4 * i + 1]);
        accW4[4 * i + 2] = fmaf(d, x.z, accW4[4 * i + 2]); accW4[4 * i + 3] = fmaf(d, x.w, accW4[4 * i + 3]);
      }
      if (!XG && tid < 32) sAcc[aB4 + tid] += sDH[tid];
      const int k = tid >> 1, hf = tid & 1;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s = fmaf(w4c[i], sDH[hf * 16 + i], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (hf == 0) {
        const int co = k >> 4, pp = k & 15, idx = sI3[pp * 16 + co];
        if (idx < 4) {
          const int y = 2 * (pp >> 2) + (idx >> 1), x = 2 * (pp & 3) + (idx & 1);
          sm[oDY3 + (co >> 2) * PL3 + ((y + 1) * P3 + x + 1) * 4 + (co & 3)] = tf32r(s * sm[oM3 + pp * 16 + co]);
        }
      }
    }
    __syncthreads();
    CF_MARK(10);
    // ================= B3: features.10 weight gradient (warps 0-9) || input gradient -> dY2 (warps 10-13)
    if (!XG && warp < 10) {
      const int mt = warp % 5, nt = warp / 5;
      const int ta = 2 * mt, tb = mt < 4 ? 2 * mt + 1 : 8;
      const int offA = (g >> 2) * PL3 + ((ta / 3) * P3 + ta % 3) * 4 + (g & 3);
      const int offB = (g >> 2) * PL3 + ((tb / 3) * P3 + tb % 3) * 4 + (g & 3);
      const int offY = (2 * nt + (g >> 2)) * PL3 + (P3 + 1) * 4 + (g & 3);
      const float ones = g == 0 ? 1.f : 0.f;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        const int base = (y * P3 + t) * 4;
        uint32_t a[4];
        a[0] = __float_as_uint(sm[oE2 + offA + base]);
        a[2] = __float_as_uint(sm[oE2 + offA + base + 16]);
        a[1] = mt < 4 ? __float_as_uint(sm[oE2 + offB + base]) : __float_as_uint(ones);
        a[3] = mt < 4 ? __float_as_uint(sm[oE2 + offB + base + 16]) : __float_as_uint(ones);
        mma_tf32(acc, a, __float_as_uint(sm[oDY3 + offY + base]), __float_as_uint(sm[oDY3 + offY + base + 16]));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int co = nt * 8 + 2 * t + (q & 1), tap = 2 * mt + (q >> 1);
        if (tap < 9) sAcc[aW3 + (co * 8 + g) * 9 + tap] += acc[q];
        else if (g == 0) sAcc[aB3 + co] += acc[q];
      }
    } else if (warp >= 10 && warp < 14) {
      const int mt = warp - 10;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t aA = smb + (oDY3 + (lj >> 1) * PL3 + ((2 * mt + (lj & 1)) * P3 + lr) * 4) * 4;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t a[4];
          ldsm4(a, aA + (ks * 2 * PL3 + ((tap / 3) * P3 + tap % 3) * 4) * 4);
          const float2 w = sWf[(wL3d >> 1) + (tap * 2 + ks) * 32 + lane];
          mma_tf32(acc, a, __float_as_uint(w.x), __float_as_uint(w.y));
        }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int y = 2 * mt + (q >> 1), ci = 2 * t + (q & 1), pq = (y * 8 + g) * 8 + ci, idx = sI2[pq];
        if (idx < 4)
          sm[oDY2 + (ci >> 2) * PL2 + ((2 * y + (idx >> 1) + 1) * P2 + 2 * g + (idx & 1) + 1) * 4 + (ci & 3)] =
              tf32r(acc[q] * sm[oM2 + pq]);
      }
    }
    __syncthreads();
    CF_MARK(11);
    // ================= B2: features.6 input gradient -> dY1 (warps 0-7) || weight gradient (warps 8-15, registers)
    if (warp >= 8) {
      if (!XG) wgrad8<16, P2, PL2>(accW, sm + oE1, sm + oDY2, (warp - 8) * 4, 4, g, t);
    } else {
      const int r0 = warp * 2;
      float2 w[3][3];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3] = sWf[(wL2d >> 1) + s * 32 + lane];
      const uint32_t aA = smb + (oDY2 + (lj >> 1) * PL2 + (r0 * P2) * 4 + ldoff8) * 4;
      slide_rows<2, 3>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (i * P2 + kx) * 16);
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int y = r0 + e + r, x = g + 8 * (q >> 1), ci = 2 * t + (q & 1), idx = sI1[(y * 16 + x) * 8 + ci];
                if (idx < 4)
                  sm[oDY1 + (ci >> 2) * PL1 + ((2 * y + (idx >> 1) + 1) * P1 + 2 * x + (idx & 1) + 1) * 4 + (ci & 3)] =
                      tf32r(r ? bot[q] : top[q]);
              }
          });
    }
    __syncthreads();
    CF_MARK(12);
    // ================= B1: features.3 weight gradient (warps 0-7, registers) || input gradient -> dE0 + arg-max tags (8-15)
    if (warp < 8) {
      if (!XG) wgrad8<32, P1, PL1>(accW, sm + oE0, sm + oDY1, warp * 16, 16, g, t);
    } else {
      const int x0 = (warp & 1) * 16, r0 = ((warp - 8) >> 1) * 8;
      float2 w[3][3];
#pragma unroll
      for (int s = 0; s < 9; ++s) w[s / 3][s % 3] = sWf[(wL1d >> 1) + s * 32 + lane];
      const uint32_t aA = smb + (oDY1 + (lj >> 1) * PL1 + (r0 * P1 + x0) * 4 + ldoff8) * 4;
      uint32_t* dD = reinterpret_cast<uint32_t*>(sm + oDE0) + (r0 * 32 + x0 + g) * 8 + 2 * t;
      const uint8_t* dI = sI0 + (r0 * 32 + x0 + g) * 8 + 2 * t;
      slide_rows<8, 3>(
          w,
          [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (i * P1 + kx) * 16);
          },
          [&](int e, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int o = ((e + r) * 32 + 8 * h) * 8;
                const uint32_t ib = *reinterpret_cast<const uint16_t*>(dI + o);      // arg-max bytes of channels 2t, 2t+1
                const float v0 = r ? bot[2 * h] : top[2 * h], v1 = r ? bot[2 * h + 1] : top[2 * h + 1];
                const uint32_t i0 = ib & 0xff, i1 = ib >> 8;
                // the 2x2 window position of the max rides in the two lowest mantissa bits (TF32 ignores them)
                uint2 st;
                st.x = i0 < 4 ? ((f2tf32(v0) & ~3u) | i0) : 0u;
                st.y = i1 < 4 ? ((f2tf32(v1) & ~3u) | i1) : 0u;
                *reinterpret_cast<uint2*>(dD + o) = st;
                bsum[0] += i0 < 4 ? v0 : 0.f;                                         // features.0 bias gradient
                bsum[1] += i1 < 4 ? v1 : 0.f;
              }
          });
    }
    __syncthreads();
    CF_MARK(13);
    if (XG) {
      // ================= features.0 input gradient: dX[y,x,ci] = sum_{tap',co} dY0[y + ky' - 1, x + kx' - 1, co] * W0[co][ci][8 - tap'].
      // dY0 (64x64x8, one non-zero per pooled window and channel) is expanded from the tagged pooled gradient into haloed
      // half-planes, 34 rows at a time (region A; e0 and dY1 are dead), and convolved like any other layer: 2 bands x 32 rows
      if (!BL && n + (int)gridDim.x < p.B) {
        prefetch_frame(p, n + gridDim.x, sm, tid);
        if (p.rng_state) draw_masks(p, n + gridDim.x, rng_call, sm, tid);
      }
      constexpr int PLD = 34 * P0 * 4;
      const uint32_t* sDE0 = reinterpret_cast<const uint32_t*>(sm + oDE0);
      const float2* wd = reinterpret_cast<const float2*>(sm + oAcc);
      float2 w[3][3];
#pragma unroll
      for (int q = 0; q < 9; ++q) w[q / 3][q % 3] = wd[q * 32 + lane];
      for (int band = 0; band < 2; ++band) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = tid; e < 2 * PLD / 4; e += NT) reinterpret_cast<float4*>(sm + oA)[e] = z4;
        __syncthreads();
        for (int e = tid; e < 18 * 256; e += NT) {          // pooled rows 16*band-1 .. 16*band+16 touch band rows 0..33
          const int py = 16 * band - 1 + (e >> 8), px = (e >> 3) & 31, co = e & 7;
          if (py < 0 || py > 31) continue;
          const uint32_t bits = sDE0[(py * 32 + px) * 8 + co];
          if (bits == 0u) continue;
          const int rb = 2 * py + (int)((bits >> 1) & 1u) - (32 * band - 1), x = 2 * px + (int)(bits & 1u);
          if (rb < 0 || rb > 33) continue;
          sm[oA + (co >> 2) * PLD + (rb * P0 + x + 1) * 4 + (co & 3)] = __uint_as_float(bits);
        }
        __syncthreads();
        {
          const int x0 = (warp & 3) * 16, r0 = (warp >> 2) * 8;
          const uint32_t aA = smb + (oA + (lj >> 1) * PLD + (r0 * P0 + x0) * 4 + ldoff8) * 4;
          float* dO = BL ? nullptr : p.dx + ((size_t)n * 4096 + (32 * band + r0) * 64 + x0 + g) * 3;
          slide_rows<8, 3>(
              w,
              [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (i * P0 + kx) * 16);
              },
              [&](int e, const float(&top)[4], const float(&bot)[4]) {
                if (BL) {
                  // d loss / d Z = sum_c dX[c] * (B - A)[c] (replaced) or (A - B)[c] (injected), main.py:395,406: lanes t = 0
                  // (channels 0, 1) and t = 1 (channel 2) of a pixel are neighbours; the differences come from the raw bytes
#pragma unroll
                  for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                      const int y = 32 * band + r0 + e + r, x = x0 + g + 8 * h;
                      const uint8_t* pa = sA8 + (y * 64 + ((x + roll) & 63)) * 3;
                      const uint8_t* pb = sB8 + (y * 64 + x) * 3;
                      float s = 0.f;
                      if (t == 0) s = (r ? bot[2 * h] : top[2 * h]) * (float)((int)pb[0] - (int)pa[0]) +
                                      (r ? bot[2 * h + 1] : top[2 * h + 1]) * (float)((int)pb[1] - (int)pa[1]);
                      else if (t == 1) s = (r ? bot[2 * h] : top[2 * h]) * (float)((int)pb[2] - (int)pa[2]);
                      s += __shfl_xor_sync(0xffffffffu, s, 1);
                      if (t == 0) {
                        float* dzp = p.dz + (size_t)n * 4096 + y * 64 + x;
                        s *= which ? -(1.f / 255.f) : (1.f / 255.f);
                        if (which == 0) {                    // first pass writes, with the regulariser's gradient (main.py:415-429)
                          const float vf = p.vpred ? 1.f - __ldg(p.vpred + n) : 1.f;
                          const float u = vf * __ldg(p.zmask + (size_t)n * 4096 + y * 64 + x);
                          reg1 += fabsf(u); reg2 = fmaf(u, u, reg2);
                          const float sg = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
                          *dzp = s + p.reg_scale * vf * (p.l1 * sg + 2.f * p.l2 * u);
                        } else {
                          *dzp += s;                         // same thread wrote it in pass 0
                        }
                      }
                    }
                } else
                if (t < 2) {                                 // columns 0..2 of the 8-wide tile are the three input channels
#pragma unroll
                  for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                      float* d = dO + ((e + r) * 64 + 8 * h) * 3 + 2 * t;
                      d[0] = r ? bot[2 * h] : top[2 * h];
                      if (t == 0) d[1] = r ? bot[2 * h + 1] : top[2 * h + 1];
                    }
                }
              });
        }
        __syncthreads();
      }
    } else {
      // ================= B0a: frame again (region A is free now), then start fetching the next frame
      stage_frame(reinterpret_cast<const uint8_t*>(sm + oU8), sm + oXB, roll, tid);
      __syncthreads();
      if (n + (int)gridDim.x < p.B) {
        prefetch_frame(p, n + gridDim.x, sm, tid);
        if (p.rng_state) draw_masks(p, n + gridDim.x, rng_call, sm, tid);
      }
      CF_MARK(14);
      // ================= B0: features.0 weight gradient, K = 4096 pixels over 16 warps (4 rows each), registers
      {
        const uint32_t* sDE0 = reinterpret_cast<const uint32_t*>(sm + oDE0);
        const float* sXb = sm + oXB;
#pragma unroll 1
        for (int yy = 0; yy < 4; ++yy) {
          const int y = warp * 4 + yy;
          const float* xa = sXb + (y * P0 + t) * 4;
          const uint32_t* de = sDE0 + ((y >> 1) * 32 + (t >> 1)) * 8 + g;
          const uint32_t pos = ((y & 1) << 1) | (t & 1);
#pragma unroll
          for (int xs = 0; xs < 8; ++xs) {
            uint32_t b0 = de[xs * 32], b1 = de[xs * 32 + 16];
            b0 = (b0 & 3u) == pos ? b0 : 0u;
            b1 = (b1 & 3u) == pos ? b1 : 0u;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              uint32_t a[4];
              a[0] = __float_as_uint(xa[offA0[mt][0] + xs * 32]);
              a[1] = __float_as_uint(xa[offA0[mt][1] + xs * 32]);
              a[2] = __float_as_uint(xa[offA0[mt][0] + xs * 32 + 16]);
              a[3] = __float_as_uint(xa[offA0[mt][1] + xs * 32 + 16]);
              mma_tf32(acc0[mt], a, b0, b1);
            }
          }
        }
      }
    }
    CF_MARK(15);
    if (trace && fr < 3) ++fr;
    // the next iteration's F0a barrier (or the one below) separates B0's reads from the next writes
  }

  if (ctat && threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); ctat[1] = gt; }
  // ---- end of the CTA's frames: combine the warps' register accumulators through shared memory (region A is free),
  //      then write this CTA's gradient: one coalesced partial vector (summed by the Adam kernel), or REDs
  asm volatile("cp.async.wait_all;\n" ::);
  __syncthreads();
  if (!XG) {
    {
      float* scr = sm + oA;                           // [16 warps][584] conv tiles, then [16 warps][216] for features.0
      wgrad8_store(accW, scr + warp * 584, g, t);
      float* scr0 = sm + oA + 16 * 584 + warp * 216;
  #pragma unroll
      for (int mt = 0; mt < 2; ++mt)
  #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int G = 2 * mt + (q >> 1), co = 2 * t + (q & 1);
          int ky, kx, ci;
          if (G < 3) {
            const int c = g < 6 + G ? g : g + 1;
            ky = G; kx = c / 3; ci = c - kx * 3;
          } else {
            ky = g; kx = 2; ci = g;
          }
          if (G < 3 || g < 3) scr0[co * 27 + ci * 9 + ky * 3 + kx] = acc0[mt][q];
        }
      // features.0 bias: lanes with equal t hold the same channels; fold the 8 g's, one value per warp and channel
      bsum[0] += __shfl_xor_sync(0xffffffffu, bsum[0], 4); bsum[1] += __shfl_xor_sync(0xffffffffu, bsum[1], 4);
      bsum[0] += __shfl_xor_sync(0xffffffffu, bsum[0], 8); bsum[1] += __shfl_xor_sync(0xffffffffu, bsum[1], 8);
      bsum[0] += __shfl_xor_sync(0xffffffffu, bsum[0], 16); bsum[1] += __shfl_xor_sync(0xffffffffu, bsum[1], 16);
      if (g == 0) { sm[oA + 16 * 800 + warp * 8 + 2 * t] = bsum[0]; sm[oA + 16 * 800 + warp * 8 + 2 * t + 1] = bsum[1]; }
    }
    __syncthreads();
    for (int e = tid; e < 584; e += NT) {
      float s1 = 0.f, s2 = 0.f;
  #pragma unroll
      for (int w = 0; w < 8; ++w) { s1 += sm[oA + w * 584 + e]; s2 += sm[oA + (w + 8) * 584 + e]; }
      sAcc[aW1 + e] = s1;                             // aB1 == aW1 + 576, aB2 == aW2 + 576
      sAcc[aW2 + e] = s2;
    }
    for (int e = tid; e < 224; e += NT) {            // features.0 weight (216) + bias (8: aB0 == aW0 + 216), fixed order
      float s0 = 0.f;
      if (e < 216) {
  #pragma unroll
        for (int w = 0; w < 16; ++w) s0 += sm[oA + 16 * 584 + w * 216 + e];
      } else {
  #pragma unroll
        for (int w = 8; w < 16; ++w) s0 += sm[oA + 16 * 800 + w * 8 + e - 216];
      }
      sAcc[aW0 + e] = s0;
    }
    __syncthreads();
    CF_MARK(17);
    grad_handover(p, sAcc, accW4, tid);
  }
  if (tid == 0) {
    // loss: one atomic per CTA into the scalar the host zeroed - or, with the grid barrier below, a slot of this CTA's
    // partial vector that CTA 0 sums in a fixed order afterwards (no memset node, bit-reproducible)
    if (p.adam_p) p.partials[(size_t)blockIdx.x * PSTRIDE + NGRAD] = loss_acc * p.inv_n;
    else if (!FWD) atomicAdd(p.loss, loss_acc * p.inv_n);
    if (BL && npass > 1) atomicAdd(p.loss + 1, loss_acc2 * p.inv_n);
  }
  if (BL) {                                          // regulariser terms: one atomic pair per warp
    reg1 = warp_sum(reg1); reg2 = warp_sum(reg2);
    if (lane == 0 && p.l1 != 0.f) atomicAdd(p.loss + 2, reg1 * p.l1 * p.inv_n * (1.f / 4096.f));
    if (lane == 0 && p.l2 != 0.f) atomicAdd(p.loss + 3, reg2 * p.l2 * p.inv_n * (1.f / 4096.f));
  }
  if (!XG && p.adam_p) adam_tail(p, sm + oA, bar_gen, tid, warp, lane, [&](int k) { CF_MARK(k); });
  if (p.rng_state && tid == 0) {                  // last CTA to finish advances the call counter (every CTA has read it)
    __threadfence();
    if (atomicAdd(&p.rng_state[1], 1ull) == gridDim.x - 1) {
      p.rng_state[1] = 0;
      p.rng_state[0] = rng_call + npass;
    }
  }
  CF_MARK(16);
  if (ctat && threadIdx.x == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); ctat[2] = gt; }
}

}  // namespace cf
}  // namespace cgs

using namespace cgs;

// Debug: point the phase trace at a device buffer of 4*24 + 4*grid int64 (NULL disables).  Not part of the product API.
extern "C" int cgs_critic_fused_set_trace(long long* dev_buf) {
  return cudaMemcpyToSymbol(cf::g_cf_trace, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -2;
}

extern "C" int cgs_critic_fused_supported(int32_t C0, int32_t C1, int32_t C2, int32_t C3, int32_t NB) {
  return C0 == 8 && C1 == 8 && C2 == 8 && C3 == 16 && NB == 32;
}

static int cf_sms() { return device_sms(); }

// CTAs the whole-step kernel runs for a batch of B frames (= rows of the partial-gradient buffer): frames per CTA are
// equalised so that no CTA idles a whole frame.
extern "C" int cgs_critic_fused_grid(int32_t B) {
  if (B <= 0) return 0;
  const int sms = cf_sms(), per = (B + sms - 1) / sms;
  return (B + per - 1) / per;
}
extern "C" int cgs_critic_fused_partial_stride(void) { return cf::PSTRIDE; }

extern "C" int cgs_critic_train_fused(const uint8_t* frames, const float* target, int32_t B, int32_t roll,
                                      const int32_t* roll_dev, const float* m_e2, const float* m_e3, const float* m_v,
                                      float p_drop, uint64_t seed, uint64_t* rng_state,
                                      const cgs_critic_weights* w, const cgs_critic_weights* g, float* partials,
                                      const cgs_adam_args* adam, float loss_grad, int32_t bce, float* pred, float* loss,
                                      void* stream) {
  CGS_REQUIRE(frames && target && w && pred && loss && B > 0 && (g || partials), "critic_train_fused: bad args");
  CGS_REQUIRE(((uintptr_t)frames & 15) == 0, "critic_train_fused: frames must be 16-byte aligned");
  CGS_REQUIRE((m_e2 != nullptr) == (m_e3 != nullptr) && (m_e2 != nullptr) == (m_v != nullptr),
              "critic_train_fused: dropout masks are all-or-none");
  CGS_REQUIRE((((uintptr_t)m_e2 | (uintptr_t)m_e3 | (uintptr_t)m_v | (uintptr_t)partials) & 15) == 0,
              "critic_train_fused: masks and partials must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cf::Params p;
  p.frames = frames; p.xin = nullptr; p.dx = nullptr; p.target = target; p.m2 = m_e2; p.m3 = m_e3; p.mv = m_v;
  p.w0 = w->w0; p.b0 = w->b0; p.w1 = w->w1; p.b1 = w->b1; p.w2 = w->w2; p.b2 = w->b2; p.w3 = w->w3; p.b3 = w->b3;
  p.w4 = w->w4; p.b4 = w->b4; p.wl1 = w->wl1; p.bl1 = w->bl1; p.wl2 = w->wl2; p.bl2 = w->bl2;
  for (int i = 0; i < 14; ++i) p.gseg[i] = nullptr;
  if (!partials) {
    float* gs[14] = {g->w0, g->b0, g->w1, g->b1, g->w2, g->b2, g->w3, g->b3, g->b4, g->wl1, g->bl1, g->wl2, g->bl2, g->w4};
    for (int i = 0; i < 14; ++i) {
      CGS_REQUIRE(gs[i] != nullptr, "critic_train_fused: gradient tensor %d is NULL", i);
      p.gseg[i] = gs[i];
    }
  }
  CGS_REQUIRE(!(rng_state && m_e2), "critic_train_fused: pass dropout masks OR an rng state, not both");
  CGS_REQUIRE(!rng_state || (p_drop > 0.f && p_drop < 1.f), "critic_train_fused: rng dropout needs 0 < p < 1");
  p.seed = seed; p.rng_state = (unsigned long long*)rng_state; p.p_drop = p_drop;
  p.keep = rng_state ? 1.f / (1.f - p_drop) : 1.f;
  p.world = 1; p.rank = 0; p.npad = 0;
  for (int r = 0; r < 16; ++r) p.ll_peer[r] = nullptr;
  p.adam_p = p.adam_g = p.adam_m = p.adam_v = nullptr;
  p.step_state = nullptr; p.bar = nullptr; p.lr = p.beta1 = p.beta2 = p.eps = 0;
  if (adam) {
    CGS_REQUIRE(partials && adam->p && adam->g && adam->m && adam->v && adam->step_state && adam->barrier,
                "critic_train_fused: in-kernel Adam needs the partial buffer and all optimizer pointers");
    CGS_REQUIRE(cgs_critic_fused_grid(B) <= 152, "critic_train_fused: grid too large for the in-kernel reduction");
    p.adam_p = adam->p; p.adam_g = adam->g; p.adam_m = adam->m; p.adam_v = adam->v;
    p.lr = adam->lr; p.beta1 = adam->beta1; p.beta2 = adam->beta2; p.eps = adam->eps;
    p.step_state = adam->step_state; p.bar = adam->barrier;
    if (adam->world > 1) {
      CGS_REQUIRE(adam->world <= 16 && adam->rank >= 0 && adam->rank < adam->world && adam->peer_recv && adam->npad >= cf::NGRAD,
                  "critic_train_fused: bad peer-memory arguments (world %d rank %d)", adam->world, adam->rank);
      p.world = adam->world; p.rank = adam->rank; p.npad = adam->npad;
      for (int r = 0; r < adam->world; ++r) p.ll_peer[r] = reinterpret_cast<unsigned long long*>(adam->peer_recv[r]);
    }
  }
  p.partials = partials;
  p.pred = pred; p.loss = loss; p.roll_dev = roll_dev; p.B = B; p.roll = roll; p.bce = bce;
  p.inv_n = 1.f / (float)B;
  p.gscale = loss_grad / (float)B;
  cudaFuncSetAttribute(cf::critic_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cf::SMEM_FLOATS * 4);
  if (!adam && cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) return check_launch("critic_train_fused.memset");
  const int grid = cgs_critic_fused_grid(B);
  if (adam) {
    // the in-kernel grid barrier needs every CTA resident at once: a cooperative launch makes the driver guarantee it
    // (or fail the launch) instead of relying on "one CTA per SM and grid <= SMs"
    void* args[] = {(void*)&p};
    const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)cf::critic_fused_kernel<0>, dim3(grid), dim3(cf::NT), args,
                                                       (size_t)cf::SMEM_FLOATS * 4, st);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      set_error("critic_train_fused: cooperative launch of %d CTAs failed: %s", grid, cudaGetErrorString(ce));
      return CGS_ECUDA;
    }
  } else {
    cf::critic_fused_kernel<0><<<grid, cf::NT, cf::SMEM_FLOATS * 4, st>>>(p);
  }
  return check_launch("critic_train_fused");
}

extern "C" int cgs_critic_loss_xgrad(const float* x, const float* target, int32_t B, const float* m_e2, const float* m_e3,
                                     const float* m_v, float p_drop, uint64_t seed, uint64_t* rng_state,
                                     const cgs_critic_weights* w, float loss_grad, int32_t bce, float* pred, float* loss, float* dx,
                                     void* stream) {
  CGS_REQUIRE(x && w && pred && B > 0 && (!dx || (target && loss)), "critic_loss_xgrad: bad args");
  CGS_REQUIRE((m_e2 != nullptr) == (m_e3 != nullptr) && (m_e2 != nullptr) == (m_v != nullptr),
              "critic_loss_xgrad: dropout masks are all-or-none");
  CGS_REQUIRE((((uintptr_t)m_e2 | (uintptr_t)m_e3 | (uintptr_t)m_v) & 15) == 0, "critic_loss_xgrad: masks must be 16-byte aligned");
  CGS_REQUIRE(!(rng_state && m_e2), "critic_loss_xgrad: pass dropout masks OR an rng state, not both");
  CGS_REQUIRE(!rng_state || (p_drop > 0.f && p_drop < 1.f), "critic_loss_xgrad: rng dropout needs 0 < p < 1");
  cudaStream_t st = (cudaStream_t)stream;
  cf::Params p;
  memset(&p, 0, sizeof(p));
  p.xin = x; p.dx = dx; p.target = target; p.m2 = m_e2; p.m3 = m_e3; p.mv = m_v;
  p.w0 = w->w0; p.b0 = w->b0; p.w1 = w->w1; p.b1 = w->b1; p.w2 = w->w2; p.b2 = w->b2; p.w3 = w->w3; p.b3 = w->b3;
  p.w4 = w->w4; p.b4 = w->b4; p.wl1 = w->wl1; p.bl1 = w->bl1; p.wl2 = w->wl2; p.bl2 = w->bl2;
  p.seed = seed; p.rng_state = (unsigned long long*)rng_state; p.p_drop = p_drop;
  p.keep = rng_state ? 1.f / (1.f - p_drop) : 1.f;
  p.world = 1;
  p.pred = pred; p.loss = loss; p.B = B; p.bce = bce;
  p.inv_n = 1.f / (float)B;
  p.gscale = loss_grad / (float)B;
  cudaFuncSetAttribute(cf::critic_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cf::SMEM_FLOATS * 4);
  if (!dx) {                                         // forward only: pred
    cudaFuncSetAttribute(cf::critic_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cf::SMEM_FLOATS * 4);
    p.target = pred;                                 // any readable [B] floats: the loss of this mode is never used
    cf::critic_fused_kernel<2><<<cgs_critic_fused_grid(B), cf::NT, cf::SMEM_FLOATS * 4, st>>>(p);
    return check_launch("critic_forward_fused");
  }
  if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) return check_launch("critic_loss_xgrad.memset");
  cf::critic_fused_kernel<1><<<cgs_critic_fused_grid(B), cf::NT, cf::SMEM_FLOATS * 4, st>>>(p);
  return check_launch("critic_loss_xgrad");
}

// pred = critic(frames) for raw uint8 frames, forward only (MODE 2 with the uint8 staging of the training kernel):
// `negpred = critic(B)` of the Hourglass loop (main.py:365-367) and extract_contrastive_data (main.py:238-260) without an
// fp32 copy of the batch.
extern "C" int cgs_critic_forward_frames(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const float* m_e2,
                                         const float* m_e3, const float* m_v, float p_drop, uint64_t seed, uint64_t* rng_state,
                                         const cgs_critic_weights* w, float* pred, void* stream) {
  CGS_REQUIRE(frames && w && pred && B > 0, "critic_forward_frames: bad args");
  CGS_REQUIRE(((uintptr_t)frames & 15) == 0, "critic_forward_frames: frames must be 16-byte aligned");
  CGS_REQUIRE((m_e2 != nullptr) == (m_e3 != nullptr) && (m_e2 != nullptr) == (m_v != nullptr),
              "critic_forward_frames: dropout masks are all-or-none");
  CGS_REQUIRE((((uintptr_t)m_e2 | (uintptr_t)m_e3 | (uintptr_t)m_v) & 15) == 0, "critic_forward_frames: masks must be 16-byte aligned");
  CGS_REQUIRE(!(rng_state && m_e2), "critic_forward_frames: pass dropout masks OR an rng state, not both");
  CGS_REQUIRE(!rng_state || (p_drop > 0.f && p_drop < 1.f), "critic_forward_frames: rng dropout needs 0 < p < 1");
  cf::Params p;
  memset(&p, 0, sizeof(p));
  p.frames = frames; p.target = pred; p.m2 = m_e2; p.m3 = m_e3; p.mv = m_v;
  p.w0 = w->w0; p.b0 = w->b0; p.w1 = w->w1; p.b1 = w->b1; p.w2 = w->w2; p.b2 = w->b2; p.w3 = w->w3; p.b3 = w->b3;
  p.w4 = w->w4; p.b4 = w->b4; p.wl1 = w->wl1; p.bl1 = w->bl1; p.wl2 = w->wl2; p.bl2 = w->bl2;
  p.seed = seed; p.rng_state = (unsigned long long*)rng_state; p.p_drop = p_drop;
  p.keep = rng_state ? 1.f / (1.f - p_drop) : 1.f;
  p.world = 1; p.pred = pred; p.B = B; p.roll = roll; p.roll_dev = roll_dev;
  p.inv_n = 1.f / (float)B; p.gscale = 0.f;
  cudaFuncSetAttribute(cf::critic_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cf::SMEM_FLOATS * 4);
  cf::critic_fused_kernel<2><<<cgs_critic_fused_grid(B), cf::NT, cf::SMEM_FLOATS * 4, (cudaStream_t)stream>>>(p);
  return check_launch("critic_forward_frames");
}

extern "C" int cgs_hg_score(const uint8_t* frames_a, const uint8_t* frames_b, int32_t B, int32_t roll, const int32_t* roll_dev,
                            const float* z, const float* target_replace, const float* target_inject,
                            const float* m_e2, const float* m_e3, const float* m_v,
                            const float* m_e2_inj, const float* m_e3_inj, const float* m_v_inj,
                            float p_drop, uint64_t seed, uint64_t* rng_state, const cgs_critic_weights* w, float loss_grad,
                            const float* vpred, float l1, float l2, float* pred_replace, float* pred_inject, float* losses,
                            float* dz, void* stream) {
  CGS_REQUIRE(frames_a && frames_b && z && target_replace && w && pred_replace && losses && dz && B > 0, "hg_score: bad args");
  CGS_REQUIRE(!target_inject || pred_inject, "hg_score: the inject pass needs pred_inject");
  CGS_REQUIRE((((uintptr_t)frames_a | (uintptr_t)frames_b) & 15) == 0, "hg_score: frames must be 16-byte aligned");
  CGS_REQUIRE((m_e2 != nullptr) == (m_e3 != nullptr) && (m_e2 != nullptr) == (m_v != nullptr), "hg_score: dropout masks are all-or-none");
  CGS_REQUIRE(!m_e2 || !target_inject || (m_e2_inj && m_e3_inj && m_v_inj), "hg_score: forced masks need a second set for the inject pass");
  CGS_REQUIRE((((uintptr_t)m_e2 | (uintptr_t)m_e3 | (uintptr_t)m_v | (uintptr_t)m_e2_inj | (uintptr_t)m_e3_inj | (uintptr_t)m_v_inj) & 15) == 0,
              "hg_score: masks must be 16-byte aligned");
  CGS_REQUIRE(!(rng_state && m_e2), "hg_score: pass dropout masks OR an rng state, not both");
  CGS_REQUIRE(!rng_state || (p_drop > 0.f && p_drop < 1.f), "hg_score: rng dropout needs 0 < p < 1");
  cudaStream_t st = (cudaStream_t)stream;
  cf::Params p;
  memset(&p, 0, sizeof(p));
  p.frames = frames_a; p.framesB = frames_b; p.zmask = z; p.target = target_replace; p.target2 = target_inject;
  p.m2 = m_e2; p.m3 = m_e3; p.mv = m_v; p.m2b = m_e2_inj; p.m3b = m_e3_inj; p.mvb = m_v_inj;
  p.w0 = w->w0; p.b0 = w->b0; p.w1 = w->w1; p.b1 = w->b1; p.w2 = w->w2; p.b2 = w->b2; p.w3 = w->w3; p.b3 = w->b3;
  p.w4 = w->w4; p.b4 = w->b4; p.wl1 = w->wl1; p.bl1 = w->bl1; p.wl2 = w->wl2; p.bl2 = w->bl2;
  p.seed = seed; p.rng_state = (unsigned long long*)rng_state; p.p_drop = p_drop;
  p.keep = rng_state ? 1.f / (1.f - p_drop) : 1.f;
  p.world = 1;
  p.pred = pred_replace; p.pred2 = pred_inject; p.loss = losses; p.dz = dz; p.B = B; p.bce = 0;
  p.roll = roll; p.roll_dev = roll_dev;
  p.vpred = vpred; p.l1 = l1; p.l2 = l2;
  p.inv_n = 1.f / (float)B;
  p.gscale = loss_grad / (float)B;
  p.reg_scale = loss_grad / ((float)B * 4096.f);
  cudaFuncSetAttribute(cf::critic_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cf::SMEM_FLOATS * 4);
  if (cudaMemsetAsync(losses, 0, 4 * sizeof(float), st) != cudaSuccess) return check_launch("hg_score.memset");
  cf::critic_fused_kernel<3><<<cgs_critic_fused_grid(B), cf::NT, cf::SMEM_FLOATS * 4, st>>>(p);
  return check_launch("hg_score");
}
