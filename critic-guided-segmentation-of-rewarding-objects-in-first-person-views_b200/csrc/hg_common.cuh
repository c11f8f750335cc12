// Building blocks of the bf16 whole-frame Hourglass kernels (hg_forward.cu, hg_backward.cu): bf16 mma.sync / ldmatrix
// wrappers, the shared-memory "tape" layout both kernels agree on, the weight-fragment pack and the row-sliding conv loop.
//
// Layout rules (all activations bf16 in shared memory, fp32 accumulation):
//   * an 8-channel map is ONE haloed plane [H+2][W+2] of 16-byte pixels: an ldmatrix row is one pixel, so one ldmatrix.x4
//     delivers a complete m16n8k16 A fragment: 16 pixels x (8 channels of tap A | 8 channels of tap B), or
//     16 pixels x (8 channels of plane A | 8 channels of plane B) for the concatenated / 16-channel operands;
//   * nn.Upsample(nearest) is an address map on the ldmatrix row addresses (haloed fine coordinate v -> (v + 1) >> 1);
//   * the RGB frame is stored "pair-duplicated": entry x = {rgb0 of haloed pixel x, rgb0 of haloed pixel x+1}, so that the
//     two 16-byte rows (entry x, entry x+2) hold all three kx taps of a filter row: ONE k16 MMA per filter row;
//   * weight gradients are GEMMs with K = pixels: both operands come from ldmatrix.trans on the same planes
//     (A rows = the 8 channels of two taps, B columns = the 8 output channels), no gathers.
#pragma once
#include <cuda_bf16.h>
#include "fused_common.cuh"

namespace cgs {
namespace hg {
using namespace cf;

constexpr int NT = 512;
// haloed bf16 planes, 16 bytes per pixel
constexpr int P1 = 34, P2 = 18, P3 = 10, P4 = 6, PX = 66;
constexpr int PB1 = P1 * P1 * 16, PB2 = P2 * P2 * 16, PB3 = P3 * P3 * 16, PB4 = P4 * P4 * 16;
constexpr int PBX = PX * PX * 16;                  // pair-duplicated RGB frame, 66 rows x 66 entries
constexpr int PLB = 18 * PX * 16;                  // one 8-channel plane of an 18-row band of the 64x64 maps
// ---- the tape: what the forward leaves per frame for the backward (byte offsets; also the head of both smem maps)
constexpr int tE0 = 0, tO0 = tE0 + PB1, tE1 = tO0 + PB1, tO1 = tE1 + PB2, tE2 = tO1 + PB2, tO2 = tE2 + PB3;
constexpr int tC3 = tO2 + PB3;                     // 6 planes @4x4: e3 (16 ch, pre-dropout) | dec[4] output broadcast (32 ch)
constexpr int tO3 = tC3 + 6 * PB4;                 // 2 planes @4x4
constexpr int tH = tO3 + 2 * PB4;                  // h = embeds[4], 32 fp32
constexpr int TAPE = tH + 128;
static_assert(TAPE == 55296 && TAPE % 16 == 0, "tape layout");

// ---- weight fragments: pack[step][lane] = uint2 {b0, b1} of the K16 x N8 matrix of that MMA step
constexpr int F_C0 = 0, F_C1 = 3, F_C2 = 9, F_C3 = 15, F_D2 = 25, F_D1 = 43, F_D0 = 52, F_M0 = 61, F_M2 = 79, F_D3 = 88;
constexpr int B_M0D = 142, B_D0D = 151, B_D1D = 157, B_M2D = 163, B_D2D = 165, B_D3D = 175;
// critic input-gradient steps (hg_score.cu): features.10 / .6 / .3 / .0 with the rotated, in/out-swapped filters
constexpr int B_C3D = 211, B_C2D = 220, B_C1D = 226, B_C0D = 232;
// masker.2 as per-tap partial products (hg_forward.cu): K = 16 input channels, N = taps 0..7 | tap 8
constexpr int F_PT = 238, NSTEPS = 240;
constexpr int F_SMEM_STEPS = F_D3;                 // the forward kernel keeps steps [0, 88) in shared memory

// masker parameters in state_dict order (= flat gradient layout of the partial vectors)
constexpr int gD0W = 0, gD0B = 1152, gD1W = 1160, gD1B = 2312, gD2W = 2320, gD2B = 4048, gD3W = 4056, gD3B = 10968,
              gD4W = 10984, gD4B = 12008, gM0W = 12040, gM0B = 13624, gM2W = 13640, gM2B = 13784, NGRAD_M = 13785;
constexpr int PSTRIDE_M = 13824;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);          // .x (low 16 bits) = lo
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// Output rows [0, R) of one 16-pixel column strip (R even), NTL output-channel tiles of 8.  Iteration i loads the NK A
// fragments of haloed input row i once and feeds output rows i, i-1, i-2 (filter rows 0, 1, 2); epi(e, nt, top, bot) gets
// the finished rows e, e+1 of channel tile nt.
template <int R, int NK, int NTL, class LoadA, class Epi>
__device__ __forceinline__ void slide_bf(const uint2 (&w)[3][NK][NTL], LoadA&& loadA, Epi&& epi) {
  float acc[NTL][4][4];
#pragma unroll
  for (int i = 0; i < R + 2; ++i) {
    uint32_t a[NK][4];
    loadA(i, a);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oi = i - ky;
      if (oi >= 0 && oi < R) {
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
          if (ky == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[nt][oi & 3][q] = 0.f;
          }
#pragma unroll
          for (int kk = 0; kk < NK; ++kk) mma_bf16(acc[nt][oi & 3], a[kk], w[ky][kk][nt].x, w[ky][kk][nt].y);
        }
      }
    }
    if (i >= 3 && ((i - 3) & 1) == 0) {
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) epi(i - 3, nt, acc[nt][(i - 3) & 3], acc[nt][(i - 2) & 3]);
    }
  }
}

// bias + ReLU + 2x2 max-pool of two finished rows, value only.  The two lanes of an x-pair (g, g^1) split the work: even g
// finishes channel 2t, odd g channel 2t+1, for both pixel halves (g, g+8).  st(h, value): pooled pixel (x0 + g + 8h) >> 1.
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

// 2x2 window sum (nearest-upsample backward) of two finished rows e, e+1 (e even): returns in s[h*2 + j] the sum for
// pooled pixel (x0 + g + 8h) >> 1, channel 2t + j; valid on both lanes of the x-pair.
__device__ __forceinline__ void sum2x2(const float (&top)[4], const float (&bot)[4], float (&s)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float v = top[q] + bot[q];
    s[q] = v + __shfl_xor_sync(0xffffffffu, v, 4);
  }
}

// raw uint8 rows -> pair-duplicated bf16 rows.  Plane row rho (0 <= rho < nrows) holds frame row y0 + rho (zeros when that
// row is outside the frame); entry e of a row = {pixel e-1, pixel e} in frame coordinates (haloed pixel e, e+1).  The halo
// halves (entry 0 low, entry 64 high, entry 65) are never written: the caller zeroes the plane once.
__device__ __forceinline__ void stage_rows(const uint8_t* __restrict__ sU8, uint8_t* __restrict__ dst, int y0, int nrows, int roll, int tid) {
  for (int e = tid; e < nrows * 64; e += NT) {
    const int rho = e >> 6, x = e & 63, y = y0 + rho;
    uint2 v = make_uint2(0u, 0u);
    if (y >= 0 && y < 64) {
      const uint8_t* s = sU8 + (y * 64 + ((x + roll) & 63)) * 3;
      // bf16(b * fl(1/255)) == bf16(b / 255.0f) for all 256 byte values (checked exhaustively): no division needed
      constexpr float k = 1.f / 255.f;
      v.x = pack_bf16(__fmul_rn((float)s[0], k), __fmul_rn((float)s[1], k));
      v.y = pack_bf16(__fmul_rn((float)s[2], k), 0.f);
    }
    uint8_t* row = dst + (size_t)rho * (PX * 16);
    *reinterpret_cast<uint2*>(row + (x + 1) * 16) = v;        // haloed pixel x+1 = low half of entry x+1 ...
    *reinterpret_cast<uint2*>(row + x * 16 + 8) = v;          // ... and high half of entry x
  }
}

// The three dropout masks of frame n, drawn exactly as cgs_dropout_masks would fill [B*512 | B*256 | B*32] floats
// (Philox4x32-10: counter = (float4 index, call), key = seed): 200 threads, one Philox call (4 draws) each.
__device__ __forceinline__ void draw_masks3(unsigned long long seed, unsigned long long call, int B, int n, float p_drop, float keep,
                                            float* m2, float* m3, float* mv, int tid) {
  if (tid >= 200) return;
  long long vec;
  float* dst;
  if (tid < 128) { vec = (long long)n * 128 + tid; dst = m2 + tid * 4; }
  else if (tid < 192) { vec = (long long)B * 128 + (long long)n * 64 + (tid - 128); dst = m3 + (tid - 128) * 4; }
  else { vec = (long long)B * 192 + (long long)n * 8 + (tid - 192); dst = mv + (tid - 192) * 4; }
  uint32_t c[4] = {(uint32_t)vec, (uint32_t)(vec >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  float4 v;
  v.x = ((float)(c[0] >> 8) * (1.0f / 16777216.0f) >= p_drop) ? keep : 0.f;
  v.y = ((float)(c[1] >> 8) * (1.0f / 16777216.0f) >= p_drop) ? keep : 0.f;
  v.z = ((float)(c[2] >> 8) * (1.0f / 16777216.0f) >= p_drop) ? keep : 0.f;
  v.w = ((float)(c[3] >> 8) * (1.0f / 16777216.0f) >= p_drop) ? keep : 0.f;
  *reinterpret_cast<float4*>(dst) = v;
}


// masker.0 (cat(X, up(o0)) 11 -> 16) + LeakyReLU on R rows of an 18-row band, both channel tiles; band row r = mask row
// 16*band - 1 + r (rows outside the frame are stored as zeros: they are masker.2's padding).  xplane: byte offset of the
// pair-duplicated frame plane whose row 0 is haloed frame row xrow0 (xrows rows); o0 sits at tO0; w0: the 18 fragment
// steps (ky*3 + j)*2 + nt of masker.0; bm0: its bias.  Shared by the forward kernel and the backward kernel's recompute.
template <int R>
__device__ __forceinline__ void m0_rows(uint8_t* band_base, int plane_stride, uint32_t smb, uint32_t xplane, int xrow0, int xrows,
                                        const uint2* w0, const float* bm0, int ya0, int rb0, int x0, int lane) {
  const int lj = lane >> 3, lr = lane & 7, g = lane >> 2, t = lane & 3, pixoff = lr + 8 * (lj & 1), chunk = lj >> 1;
  uint2 w[3][3][2];
#pragma unroll
  for (int s = 0; s < 18; ++s) w[s / 6][(s >> 1) % 3][s & 1] = w0[s * 32 + lane];
  const int hv0 = ya0 + rb0;                         // ya0 = mask row of band row 0; haloed frame row of input row i = hv0 + i
  const uint32_t aX = smb + xplane + (uint32_t)((x0 + pixoff + 2 * chunk) * 16);
  const int cA = (x0 + pixoff + chunk + 1) >> 1, cB = (x0 + pixoff + 3) >> 1;
  float bias[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) { bias[nt][0] = bm0[nt * 8 + 2 * t]; bias[nt][1] = bm0[nt * 8 + 2 * t + 1]; }
  slide_bf<R, 3, 2>(
      w,
      [&](int i, uint32_t(&a)[3][4]) {
        const int hv = hv0 + i;
        const int xr = min(max(hv - xrow0, 0), xrows - 1);   // rows -1 / 66 only feed band rows that are stored as zeros
        ldsm4(a[0], aX + (uint32_t)(xr * (PX * 16)));
        const int sy = (hv + 1) >> 1;                 // 0 and 33 are o0's zero halo rows
        ldsm4(a[1], smb + tO0 + (uint32_t)((sy * P1 + cA) * 16));
        ldsm2(a[2][0], a[2][1], smb + tO0 + (uint32_t)((sy * P1 + cB) * 16));
        a[2][2] = a[2][3] = 0u;
      },
      [&](int e, int nt, const float(&top)[4], const float(&bot)[4]) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int rb = rb0 + e + r, ya = ya0 + rb;
          const bool inside = ya >= 0 && ya < 64;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v0 = (r ? bot[2 * h] : top[2 * h]) + bias[nt][0], v1 = (r ? bot[2 * h + 1] : top[2 * h + 1]) + bias[nt][1];
            v0 = v0 > 0.f ? v0 : v0 * kLeakySlope;
            v1 = v1 > 0.f ? v1 : v1 * kLeakySlope;
            *reinterpret_cast<uint32_t*>(band_base + nt * plane_stride + (rb * PX + x0 + g + 8 * h + 1) * 16 + 4 * t) = inside ? pack_bf16(v0, v1) : 0u;
          }
        }
      });
}

// the band's 18 rows over 16 warps: 4 column strips x 4 row segments (6 + 4 + 4 + 4)
__device__ __forceinline__ void m0_band(uint8_t* band_base, uint32_t smb, uint32_t xplane, int xrow0, int xrows, const uint2* w0,
                                        const float* bm0, int band, int warp, int lane) {
  const int x0 = (warp & 3) * 16, seg = warp >> 2;
  if (seg == 0) m0_rows<6>(band_base, PLB, smb, xplane, xrow0, xrows, w0, bm0, 16 * band - 1, 0, x0, lane);
  else m0_rows<4>(band_base, PLB, smb, xplane, xrow0, xrows, w0, bm0, 16 * band - 1, 2 + 4 * seg, x0, lane);
}

// a 10-row band (8 mask rows + one row of overlap each side; band row r = mask row 8*band - 1 + r): 4 strips x (4 + 2 + 2 + 2) rows
constexpr int PLB8 = 10 * PX * 16;
__device__ __forceinline__ void m0_band8(uint8_t* band_base, uint32_t smb, uint32_t xplane, int xrow0, int xrows, const uint2* w0,
                                         const float* bm0, int band, int warp, int lane) {
  const int x0 = (warp & 3) * 16, seg = warp >> 2;
  if (seg == 0) m0_rows<4>(band_base, PLB8, smb, xplane, xrow0, xrows, w0, bm0, 8 * band - 1, 0, x0, lane);
  else m0_rows<2>(band_base, PLB8, smb, xplane, xrow0, xrows, w0, bm0, 8 * band - 1, 2 + 2 * seg, x0, lane);
}

// debug: clock64() at phase boundaries of CTA 0's first two frames (tools/hg_trace.py); NULL = off
#define HG_MARK(k)                                                    \
  do {                                                                \
    if (trace && tid == 0 && fr < 2) trace[fr * 32 + (k)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// weight fragment pack: element (k, n) of the K16 x N8 matrix of MMA step s (0 where the slot is padding)
struct PackSrc {
  const float *w0, *w1, *w2, *w3, *d0, *d1, *d2, *d3, *m0, *m2;
};

__device__ inline float pack_wk(const PackSrc& p, int s, int k, int n) {
  if (s < F_C1) {                                   // features.0: step ky; k = kx*4 + c (pair-duplicated frame)
    const int ky = s, kx = k >> 2, c = k & 3;
    return (kx < 3 && c < 3) ? p.w0[((n * 3 + c) * 3 + ky) * 3 + kx] : 0.f;
  }
  if (s < F_C3) {                                   // features.3 / features.6: step ky*2 + h; h 0 = taps (ky,0 | ky,1), h 1 = (ky,2 | 0)
    const float* w = s < F_C2 ? p.w1 : p.w2;
    const int q = s < F_C2 ? s - F_C1 : s - F_C2, ky = q >> 1, h = q & 1;
    if (h && k >= 8) return 0.f;
    const int kx = h ? 2 : (k >> 3), ci = k & 7;
    return w[((n * 8 + ci) * 3 + ky) * 3 + kx];
  }
  if (s < F_D2) {                                   // features.10 (8 -> 16): step tp*2 + nt; taps (2tp | 2tp+1)
    const int q = s - F_C3, tp = q >> 1, nt = q & 1, tap = 2 * tp + (k >> 3), ci = k & 7;
    return tap > 8 ? 0.f : p.w3[((nt * 8 + n) * 8 + ci) * 9 + tap];
  }
  if (s < F_D1) {                                   // dec[2] (24 -> 8): step tap*2 + kc; kc 0 = channels 0..15, kc 1 = 16..23 | 0
    const int q = s - F_D2, tap = q >> 1, kc = q & 1;
    if (kc && k >= 8) return 0.f;
    return p.d2[(n * 24 + kc * 16 + k) * 9 + tap];
  }
  if (s < F_M0) {                                   // dec[1] / dec[0] (16 -> 8): step tap; k = concatenated channel
    const float* w = s < F_D0 ? p.d1 : p.d0;
    const int tap = s < F_D0 ? s - F_D1 : s - F_D0;
    return w[(n * 16 + k) * 9 + tap];
  }
  if (s < F_M2) {                                   // masker.0 (11 -> 16): step (ky*3 + j)*2 + nt; j 0 = RGB, 1 = o0 (kx 0|1), 2 = o0 (kx 2|-)
    const int q = s - F_M0, nt = q & 1, j = (q >> 1) % 3, ky = (q >> 1) / 3, co = nt * 8 + n;
    if (j == 0) {
      const int kx = k >> 2, c = k & 3;
      return (kx < 3 && c < 3) ? p.m0[((co * 11 + c) * 3 + ky) * 3 + kx] : 0.f;
    }
    if (j == 2 && k >= 8) return 0.f;
    const int kx = j == 1 ? (k >> 3) : 2, ci = 3 + (k & 7);
    return p.m0[((co * 11 + ci) * 3 + ky) * 3 + kx];
  }
  if (s < F_D3) return n == 0 ? p.m2[k * 9 + (s - F_M2)] : 0.f;        // masker.2 (16 -> 1): step tap
  if (s < B_M0D) {                                  // dec[3] (48 -> 16): step (tap*3 + kc)*2 + nt
    const int q = s - F_D3, nt = q & 1, kc = (q >> 1) % 3, tap = (q >> 1) / 3;
    return p.d3[((nt * 8 + n) * 48 + kc * 16 + k) * 9 + tap];
  }
  // ---- input-gradient (dgrad) steps: tap' runs over the haloed output gradient, the filter is rotated: tap = 8 - tap'
  if (s < B_D0D) return p.m0[(k * 11 + 3 + n) * 9 + 8 - (s - B_M0D)];   // masker.0 -> up(o0): k = co (16), n = ci
  if (s < B_M2D) {                                  // dec[0] / dec[1] -> upsampled half: step ky'*2 + h; k = (tap' pair) x co (8)
    const float* w = s < B_D1D ? p.d0 : p.d1;
    const int q = s < B_D1D ? s - B_D0D : s - B_D1D, ky = q >> 1, h = q & 1;
    if (h && k >= 8) return 0.f;
    const int kx = h ? 2 : (k >> 3), co = k & 7;
    return w[(co * 16 + 8 + n) * 9 + 8 - (ky * 3 + kx)];
  }
  if (s < B_D2D) return k < 9 ? p.m2[((s - B_M2D) * 8 + n) * 9 + 8 - k] : 0.f;     // masker.2 -> m0: k = tap', n = ci
  if (s < B_D3D) {                                  // dec[2] -> up(o3): step tp*2 + nt; k = (tap' pair) x co (8); n: ci = 8 + nt*8 + n
    const int q = s - B_D2D, nt = q & 1, tp = 2 * (q >> 1) + (k >> 3), co = k & 7;
    return tp > 8 ? 0.f : p.d2[(co * 24 + 8 + nt * 8 + n) * 9 + 8 - tp];
  }
  if (s < B_C3D) {                                  // dec[3] -> dec[4] output: step tap'*4 + nt; k = co (16); n: ci = 16 + nt*8 + n
    const int q = s - B_D3D, nt = q & 3, tp = q >> 2;
    return p.d3[(k * 48 + 16 + nt * 8 + n) * 9 + 8 - tp];
  }
  if (s < B_C2D) return p.w3[(k * 8 + n) * 9 + 8 - (s - B_C3D)];          // features.10 -> its input: k = co (16), n = ci (8)
  if (s < F_PT) {                                   // features.6 / .3 / .0 -> their inputs: step ky'*2 + h; k = (tap' pair) x co (8)
    const float* w = s < B_C1D ? p.w2 : (s < B_C0D ? p.w1 : p.w0);
    const int q = s - (s < B_C1D ? B_C2D : (s < B_C0D ? B_C1D : B_C0D)), ky = q >> 1, h = q & 1;
    if (h && k >= 8) return 0.f;
    const int kx = h ? 2 : (k >> 3), co = k & 7, tap = 8 - (ky * 3 + kx);
    if (s < B_C0D) return w[(co * 8 + n) * 9 + tap];
    return n < 3 ? w[(co * 3 + n) * 9 + tap] : 0.f;  // features.0: 3 input channels
  }
  {                                                 // masker.2, one column per filter tap: step nt; k = ci (16); n: tap = nt*8 + n
    const int tap = (s - F_PT) * 8 + n;
    return tap < 9 ? p.m2[k * 9 + tap] : 0.f;
  }
}


// One (source, kx group) weight-gradient triple over output-gradient rows [y0, y0 + R) of one 16-pixel strip:
// acc[ky] += A(input row i) x B(gradient row i - ky).  loadA(i, a): the A fragment of haloed input row i (i = y + ky);
// loadB(y, b0, b1): the B fragment (16 pixels x 8 output channels) of gradient row y.
template <int R, class LoadA, class LoadB>
__device__ __forceinline__ void wgrad_slide(float (&acc)[3][4], int y0, LoadA&& loadA, LoadB&& loadB) {
  uint32_t b[3][2];
#pragma unroll
  for (int i = 0; i < R + 2; ++i) {
    uint32_t a[4];
    loadA(y0 + i, a);
    if (i < R) loadB(y0 + i, b[i % 3][0], b[i % 3][1]);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = i - ky;
      if (y >= 0 && y < R) mma_bf16(acc[ky], a, b[y % 3][0], b[y % 3][1]);
    }
  }
}

// ---- TMA bulk copies (cp.async.bulk: one instruction moves a whole contiguous block between global and shared memory
//      through the async proxy; SASS: UBLKCP) with mbarrier completion for loads and bulk-group completion for stores
static __device__ int g_hg_timeout = 0;     // set when a bounded mbarrier wait gave up (cgs_hg_status)

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// bounded wait (never hang the device): on time-out the flag is raised and the caller carries on with whatever has landed
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return;
  }
  atomicExch(&g_hg_timeout, 1);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::); }

}  // namespace hg
}  // namespace cgs
