// The masker's backward of one segmentation_training step (reference main.py:462, the autograd backward of nets.py:494-523)
// as ONE persistent kernel, bf16 tensor-core operands with fp32 accumulation, chfak = 1, frozen critic:
//   d loss / d mask (from cgs_hg_score) -> sigmoid' -> masker.2 wgrad + dgrad -> LeakyReLU' -> masker.0 wgrad + dgrad into
//   the upsampled o0 -> 2x2 sum -> dec[0] wgrad + dgrad -> ... -> dec[3] -> dec[4] wgrad; all 13,785 masker gradients.
// Inputs per frame: the 12 KB uint8 frame, the forward's tape (cgs_hg_forward: every skip / decoder activation as haloed
// bf16 planes), the mask and its gradient.  The 16-channel 64x64 maps (masker.0's activation and its gradient) only exist
// as 18-row bands in shared memory: the activation is recomputed per band from the frame and o0.
// Weight gradients are GEMMs with K = pixels whose operands both come from ldmatrix.trans on the activation planes; a
// (source plane, kx group) "triple" of three filter-row tiles slides down the rows like the forward convolutions do.
// Accumulators live in registers (masker.0, masker.2, dec[4]) or in shared memory in MMA fragment order (dec[0..3]) over
// ALL frames of the CTA; each CTA then writes ONE partial gradient vector (state_dict order) that
// cgs_adam_step_partials sums in a fixed order: no atomics in global memory, bit-reproducible.
// Gradients w.r.t. the skip inputs (embeds) and the frame are not computed: the critic is frozen (main.py:334).
#include <string.h>
#include "hg_common.cuh"

namespace cgs {
namespace hg {

// ---- shared memory map of the backward kernel (byte offsets); the tape sits at offset 0 exactly as the forward wrote it
constexpr int kU8 = TAPE;                           // raw frame bytes
constexpr int kXB = kU8 + 12288;                    // pair-duplicated frame rows of the band: 20 rows
constexpr int kM0 = kXB + 20 * PX * 16;             // masker.0 activation band (2 planes x 18 rows) -> its gradient, in place
constexpr int DLP = 68;                             // pitch of the dlogit band (fp32): column x+1, columns 0 and 65 zero
constexpr int kDL = kM0 + 2 * PLB;                  // 20 rows
constexpr int kDO0 = kDL + 20 * DLP * 4;            // d o0 (32x32x8), d o1, d o2, d o3 (2 planes): haloed bf16 planes
constexpr int kDO1 = kDO0 + PB1, kDO2 = kDO1 + PB2, kDO3 = kDO2 + PB3;
constexpr int kW = kDO3 + 2 * PB4;                  // weight fragments: masker.0 fprop (18 steps) | B_M0D .. B_D2D (23 steps)
constexpr int W_M0F = 0, W_BWD = 18, W_STEPS = 18 + (B_D2D - B_M0D);
constexpr int kAcc = kW + W_STEPS * 256;            // fp32 accumulator tiles in MMA fragment order, 512 B each
constexpr int T_D0 = 0, T_D1 = 24, T_D2 = 36, T_D3 = 51, N_TILES = 111;
constexpr int kMisc = kAcc + N_TILES * 512;         // dD4[32] | bm0[16]
constexpr int B_SMEM = kMisc + 256;
static_assert(B_SMEM <= 227 * 1024, "backward kernel: shared memory budget");
static_assert(kU8 % 16 == 0 && kXB % 16 == 0 && kM0 % 16 == 0 && kDL % 16 == 0 && kDO0 % 16 == 0 && kW % 16 == 0 && kAcc % 16 == 0,
              "16-byte alignment");
static_assert(NGRAD_M * 4 <= TAPE, "the partial vector is assembled in the tape region");
static_assert(kDL - kXB >= TAPE, "the next frame's tape is prefetched into the frame-row + band region");

struct BwdParams {
  const uint8_t* frames;
  const uint8_t* tape;
  const float* z;
  const float* dz;
  const uint2* pack;
  const float* bm0;
  const int* roll_dev;
  int B, roll;
  float* partials;
  float* dbg;           // debug: frame 0's d o0 .. d o3 planes and d(dec[4] output), see tests/test_gpu_hg.py
};

constexpr uint32_t ONES2 = 0x3F803F80u;             // bf16 (1, 1)

// accumulate three fragment tiles held in registers into shared-memory tiles owned by this warp
__device__ __forceinline__ void flush3(float* tiles, const float (&acc)[3][4], int lane) {
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    float4* d = reinterpret_cast<float4*>(tiles + ky * 128) + lane;
    float4 v = *d;
    v.x += acc[ky][0]; v.y += acc[ky][1]; v.z += acc[ky][2]; v.w += acc[ky][3];
    *d = v;
  }
}

__device__ long long* g_hgb_trace = nullptr;
int set_fwd_trace(long long* b);
}
namespace hs { int set_trace(long long* b); }
namespace hg {

__global__ void __launch_bounds__(NT, 1) hg_backward_kernel(const BwdParams p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int lj = lane >> 3, lr = lane & 7, pixoff = lr + 8 * (lj & 1), chunk = lj >> 1;
  // ldmatrix.trans operand roles: matrix (lj & 1) = tap / plane selector, matrix (lj >> 1) = pixel half
  const int tsel = lj & 1, thalf = lj >> 1, tpix = lr + 8 * thalf;
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(smraw);
  const uint2* sWf = reinterpret_cast<const uint2*>(smraw + kW);
  float* sDL = reinterpret_cast<float*>(smraw + kDL);
  float* sAcc = reinterpret_cast<float*>(smraw + kAcc);
  float* sDD4 = reinterpret_cast<float*>(smraw + kMisc);
  float* sBm0 = sDD4 + 32;
  const float* sH = reinterpret_cast<const float*>(smraw + tH);
  const uint32_t ones = g == 0 ? ONES2 : 0u;
  long long* trace = blockIdx.x == 0 ? g_hgb_trace : nullptr;
  int fr = 0;

  // ---- prologue: zero everything once (halos and accumulator tiles stay / start zero), weight fragments, the load barrier
  __shared__ __align__(8) unsigned long long s_bar;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
  uint32_t phase = 0u;
  bool prefetched = false;
  if (tid == 0) mbar_init(bar, 1);
  {
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int e = tid; e < B_SMEM / 16; e += NT) reinterpret_cast<uint4*>(smraw)[e] = z4;
  }
  __syncthreads();
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.pack);
    for (int e = tid; e < 18 * 16; e += NT) reinterpret_cast<uint4*>(smraw + kW)[e] = __ldg(src + F_M0 * 16 + e);
    for (int e = tid; e < (B_D2D - B_M0D) * 16; e += NT) reinterpret_cast<uint4*>(smraw + kW + 18 * 256)[e] = __ldg(src + B_M0D * 16 + e);
    if (tid < 16) sBm0[tid] = __ldg(p.bm0 + tid);
  }
  int roll = p.roll_dev ? *p.roll_dev : p.roll;
  roll = ((roll % 64) + 64) & 63;

  // register accumulators over all frames of this CTA
  float accM0[3][4];      // masker.0: warps 0-11 own (triple, channel tile, row half)
  float accM2[2][4];      // masker.2: every warp holds a partial over its pixel groups
  float accD4[2] = {0.f, 0.f}, accB4 = 0.f, bsum2 = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) accM0[i >> 2][i & 3] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) accM2[i >> 2][i & 3] = 0.f;

  for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
    fence_proxy_async();                               // our generic-proxy accesses to the tape region precede the bulk writes
    __syncthreads();                                   // everybody is done with the previous frame's tape and bytes
    HG_MARK(0);
    // the frame's tape (54 KB) and raw bytes (12 KB) as two TMA bulk loads (cp.async.bulk) completing on one mbarrier.  From
    // the CTA's second frame on they were put in flight during the previous frame's decoder phases, the tape into the (then
    // dead) band region: it is moved home here, leaving the zeros the band planes' halos need
    if (!prefetched && tid == 0) {
      mbar_expect_tx(bar, TAPE + 12288);
      bulk_g2s(smb, p.tape + (size_t)n * TAPE, TAPE, bar);
      bulk_g2s(smb + kU8, p.frames + (size_t)n * 12288, 12288, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    if (prefetched) {
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      uint4* src = reinterpret_cast<uint4*>(smraw + kXB);
      uint4* dst = reinterpret_cast<uint4*>(smraw);
      for (int e = tid; e < TAPE / 16; e += NT) { dst[e] = src[e]; src[e] = z4; }
      __syncthreads();
    }
    HG_MARK(1);
    const float* zf = p.z + (size_t)n * 4096;
    const float* dzf = p.dz + (size_t)n * 4096;

    for (int band = 0; band < 4; ++band) {
      HG_MARK(2 + 5 * band);
      // ================= B1: the band's 20 frame rows (pair-duplicated bf16) and 20 rows of d logit = dZ * Z * (1 - Z)
      // (the mask / mask-gradient loads are put in flight first and consumed after the frame rows are staged)
      float zz[3], dd[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int e = tid + NT * i, rho = e >> 6, x = e & 63, y = 16 * band - 2 + rho;
        const bool ok = e < 20 * 64 && y >= 0 && y < 64;
        zz[i] = ok ? __ldg(zf + y * 64 + x) : 0.f;
        dd[i] = ok ? __ldg(dzf + y * 64 + x) : 0.f;
      }
      stage_rows(smraw + kU8, smraw + kXB, 16 * band - 2, 20, roll, tid);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int e = tid + NT * i, rho = e >> 6, x = e & 63;
        if (e < 20 * 64) {
          const float v = dd[i] * zz[i] * (1.f - zz[i]);    // 0 outside the frame
          if (rho >= 2 && rho < 18) bsum2 += v;             // masker.2 bias gradient: every mask row exactly once
          sDL[rho * DLP + x + 1] = v;
        }
      }
      __syncthreads();
      HG_MARK(3 + 5 * band);
      // ================= B2: masker.0 + LeakyReLU recomputed for the band's 18 rows
      m0_band(smraw + kM0, smb, kXB, 16 * band - 1, 20, sWf + W_M0F * 32, sBm0, band, warp, lane);
      __syncthreads();
      HG_MARK(4 + 5 * band);
      // ================= B3: masker.2 weight gradient: M = 16 input channels, N = 9 taps, K = pixels of the band
      // dW2[ci][ky][kx] += m0[r][c][ci] * dlogit[row r - ky + 1][col c - kx]  (band coordinates, interior mask rows only)
      for (int pg = warp; pg < 72; pg += 16) {
        const int r = pg >> 2, s = pg & 3;
        uint32_t a[4];
        ldsm4t(a, smb + kM0 + (uint32_t)(tsel * PLB + (r * PX + 1 + 16 * s + tpix) * 16));
        {
          const int ky = g / 3, kx = g - 3 * ky;     // first tile: taps 0..7 (n = g)
          const int rho = r - ky + 2;
          uint32_t b0 = 0u, b1 = 0u;
          if (rho >= 2 && rho < 18) {
            const float* q = sDL + rho * DLP + 16 * s + 2 * t - kx + 2;
            b0 = pack_bf16(q[0], q[1]);
            b1 = pack_bf16(q[8], q[9]);
          }
          mma_bf16(accM2[0], a, b0, b1);
        }
        {
          const int rho = r;                          // second tile: tap 8 = (2, 2) in column n = 0
          uint32_t b0 = 0u, b1 = 0u;
          if (g == 0 && rho >= 2 && rho < 18) {
            const float* q = sDL + rho * DLP + 16 * s + 2 * t;
            b0 = pack_bf16(q[0], q[1]);
            b1 = pack_bf16(q[8], q[9]);
          }
          mma_bf16(accM2[1], a, b0, b1);
        }
      }
      __syncthreads();
      HG_MARK(5 + 5 * band);
      // ================= B4: masker.2 input gradient x LeakyReLU' -> gradient of masker.0's output, in place of the activation
      for (int pg = warp; pg < 72; pg += 16) {
        const int r = pg >> 2, s = pg & 3;
        uint32_t a[4];
        {
          // A[m = pixel][k = tap']: dlogit[row r + ky'][col x + kx'] ; k = 2t, 2t+1 (and tap' 8 on t == 0)
          const int ta = 2 * t, tb = 2 * t + 1;
          const float* q0 = sDL + (r + ta / 3) * DLP + 16 * s + g + ta % 3;
          const float* q1 = sDL + (r + tb / 3) * DLP + 16 * s + g + tb % 3;
          a[0] = pack_bf16(q0[0], q1[0]);
          a[1] = pack_bf16(q0[8], q1[8]);
          const float* q8 = sDL + (r + 2) * DLP + 16 * s + g + 2;
          a[2] = t == 0 ? pack_bf16(q8[0], 0.f) : 0u;
          a[3] = t == 0 ? pack_bf16(q8[8], 0.f) : 0u;
        }
        const int ya = 16 * band - 1 + r;
        const bool inside = ya >= 0 && ya < 64;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          float c[4] = {0.f, 0.f, 0.f, 0.f};
          const uint2 w = sWf[(W_BWD + B_M2D - B_M0D + nt) * 32 + lane];
          mma_bf16(c, a, w.x, w.y);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t* q = reinterpret_cast<uint32_t*>(smraw + kM0 + nt * PLB + (r * PX + 16 * s + g + 8 * h + 1) * 16 + 4 * t);
            const uint32_t m = *q;
            const float d0 = bf16_lo(m) > 0.f ? c[2 * h] : c[2 * h] * kLeakySlope;
            const float d1 = bf16_hi(m) > 0.f ? c[2 * h + 1] : c[2 * h + 1] * kLeakySlope;
            *q = inside ? pack_bf16(d0, d1) : 0u;
          }
        }
      }
      __syncthreads();
      HG_MARK(6 + 5 * band);
      // ================= B5: masker.0 weight gradient (warps 0-11, registers) || input gradient into up(o0), 2x2 sum (12-15)
      if (warp < 12) {
        const int tr = warp % 3, nt = (warp / 3) & 1, kh = warp / 6;
        const int r0 = 1 + 8 * kh;                    // gradient rows r0 .. r0+7 of the band = mask rows 16*band + 8*kh ..
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          const int x0 = 16 * s;
          auto loadB = [&](int r, uint32_t& b0, uint32_t& b1) {
            ldsm2t(b0, b1, smb + kM0 + (uint32_t)(nt * PLB + (r * PX + 1 + x0 + lr + 8 * (lj & 1)) * 16));
          };
          // input row i of the triple = band row r + ky - 1 -> frame-plane row rho = r + ky (rows r0 .. r0+9 for ky 0..2)
          if (tr == 0) {
            wgrad_slide<8>(accM0, r0,
                           [&](int i, uint32_t(&a)[4]) { ldsm4t(a, smb + kXB + (uint32_t)((i * PX + x0 + tpix + 2 * tsel) * 16)); }, loadB);
          } else if (tr == 1) {
            wgrad_slide<8>(accM0, r0,
                           [&](int i, uint32_t(&a)[4]) {
                             const int hv = 16 * band - 1 + i, sy = (hv + 1) >> 1;
                             ldsm4t(a, smb + tO0 + (uint32_t)((sy * P1 + ((x0 + tpix + tsel + 1) >> 1)) * 16));
                           },
                           loadB);
          } else {
            wgrad_slide<8>(accM0, r0,
                           [&](int i, uint32_t(&a)[4]) {
                             const int hv = 16 * band - 1 + i, sy = (hv + 1) >> 1;
                             ldsm2t(a[0], a[2], smb + tO0 + (uint32_t)((sy * P1 + ((x0 + lr + 8 * (lj & 1) + 3) >> 1)) * 16));
                             a[1] = a[3] = ones;
                           },
                           loadB);
          }
        }
      } else {
        const int x0 = (warp - 12) * 16;
        uint2 w[3][3][1];
#pragma unroll
        for (int q = 0; q < 9; ++q) w[q / 3][q % 3][0] = sWf[(W_BWD + q) * 32 + lane];
        const uint32_t aA = smb + kM0 + (uint32_t)(chunk * PLB + (x0 + pixoff) * 16);
        slide_bf<16, 3, 1>(
            w,
            [&](int i, uint32_t(&a)[3][4]) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) ldsm4(a[kx], aA + (uint32_t)((i * PX + kx) * 16));
            },
            [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
              float sm4[4];
              sum2x2(top, bot, sm4);
              if (!(g & 1)) {
                const int py = 8 * band + (e >> 1), px = (x0 + g) >> 1;
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  *reinterpret_cast<uint32_t*>(smraw + kDO0 + ((py + 1) * P1 + px + 4 * h + 1) * 16 + 4 * t) = pack_bf16(sm4[2 * h], sm4[2 * h + 1]);
              }
            });
      }
      fence_proxy_async();                             // (last band) our reads of the band region precede the bulk prefetch into it
      __syncthreads();
    }

    HG_MARK(22);
    // the band region (frame rows, masker.0 band) and the raw bytes are dead until the next frame: fetch its tape and bytes now
    prefetched = n + (int)gridDim.x < p.B;
    if (prefetched && tid == 0) {
      const int n2 = n + gridDim.x;
      mbar_expect_tx(bar, TAPE + 12288);
      bulk_g2s(smb + kXB, p.tape + (size_t)n2 * TAPE, TAPE, bar);
      bulk_g2s(smb + kU8, p.frames + (size_t)n2 * 12288, 12288, bar);
    }
    // ================= D0: dec[0] (cat(e0, up(o1)) -> o0 on 32x32): weight gradient (warps 0-7) || input gradient -> d o1 (8-15)
    if (warp < 8) {
      const int tr = warp & 3, kh = warp >> 2, src = tr >> 1, kxg = tr & 1;
      float acc[3][4];
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[i >> 2][i & 3] = 0.f;
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        const int x0 = 16 * s;
        auto loadB = [&](int y, uint32_t& b0, uint32_t& b1) {
          ldsm2t(b0, b1, smb + kDO0 + (uint32_t)(((y + 1) * P1 + 1 + x0 + lr + 8 * (lj & 1)) * 16));
        };
        auto loadA = [&](int i, uint32_t(&a)[4]) {     // haloed input row i = y + ky
          if (kxg == 0) {
            const int v = x0 + tpix + tsel;            // haloed fine column of tap kx = tsel
            ldsm4t(a, src == 0 ? smb + tE0 + (uint32_t)((i * P1 + v) * 16)
                               : smb + tO1 + (uint32_t)((((i + 1) >> 1) * P2 + ((v + 1) >> 1)) * 16));
          } else {
            const int v = x0 + lr + 8 * (lj & 1) + 2;
            ldsm2t(a[0], a[2], src == 0 ? smb + tE0 + (uint32_t)((i * P1 + v) * 16)
                                        : smb + tO1 + (uint32_t)((((i + 1) >> 1) * P2 + ((v + 1) >> 1)) * 16));
            a[1] = a[3] = src == 0 ? ones : 0u;
          }
        };
        wgrad_slide<16>(acc, 16 * kh, loadA, loadB);
      }
      flush3(sAcc + (T_D0 + (kh * 4 + tr) * 3) * 128, acc, lane);
    } else {
      const int x0 = (warp & 1) * 16, r0 = ((warp - 8) >> 1) * 8;
      uint2 w[3][2][1];
#pragma unroll
      for (int q = 0; q < 6; ++q) w[q >> 1][q & 1][0] = sWf[(W_BWD + B_D0D - B_M0D + q) * 32 + lane];
      const uint32_t aA = smb + kDO0 + (uint32_t)((r0 * P1 + x0 + pixoff + chunk) * 16);
      const uint32_t aB = smb + kDO0 + (uint32_t)((r0 * P1 + x0 + pixoff + 2) * 16);
      slide_bf<8, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P1 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P1 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
            float sm4[4];
            sum2x2(top, bot, sm4);
            if (!(g & 1)) {
              const int py = (r0 + e) >> 1, px = (x0 + g) >> 1;
#pragma unroll
              for (int h = 0; h < 2; ++h)
                *reinterpret_cast<uint32_t*>(smraw + kDO1 + ((py + 1) * P2 + px + 4 * h + 1) * 16 + 4 * t) = pack_bf16(sm4[2 * h], sm4[2 * h + 1]);
            }
          });
    }
    __syncthreads();
    HG_MARK(23);
    // ================= D1: dec[1] (cat(e1, up(o2)) -> o1 on 16x16): weight gradient (warps 0-3) || input gradient -> d o2 (4-11)
    if (warp < 4) {
      const int tr = warp, src = tr >> 1, kxg = tr & 1;
      float acc[3][4];
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[i >> 2][i & 3] = 0.f;
      auto loadB = [&](int y, uint32_t& b0, uint32_t& b1) {
        ldsm2t(b0, b1, smb + kDO1 + (uint32_t)(((y + 1) * P2 + 1 + lr + 8 * (lj & 1)) * 16));
      };
      auto loadA = [&](int i, uint32_t(&a)[4]) {
        if (kxg == 0) {
          const int v = tpix + tsel;
          ldsm4t(a, src == 0 ? smb + tE1 + (uint32_t)((i * P2 + v) * 16)
                             : smb + tO2 + (uint32_t)((((i + 1) >> 1) * P3 + ((v + 1) >> 1)) * 16));
        } else {
          const int v = lr + 8 * (lj & 1) + 2;
          ldsm2t(a[0], a[2], src == 0 ? smb + tE1 + (uint32_t)((i * P2 + v) * 16)
                                      : smb + tO2 + (uint32_t)((((i + 1) >> 1) * P3 + ((v + 1) >> 1)) * 16));
          a[1] = a[3] = src == 0 ? ones : 0u;
        }
      };
      wgrad_slide<16>(acc, 0, loadA, loadB);
      flush3(sAcc + (T_D1 + tr * 3) * 128, acc, lane);
    } else if (warp < 12) {
      const int r0 = (warp - 4) * 2;
      uint2 w[3][2][1];
#pragma unroll
      for (int q = 0; q < 6; ++q) w[q >> 1][q & 1][0] = sWf[(W_BWD + B_D1D - B_M0D + q) * 32 + lane];
      const uint32_t aA = smb + kDO1 + (uint32_t)((r0 * P2 + pixoff + chunk) * 16);
      const uint32_t aB = smb + kDO1 + (uint32_t)((r0 * P2 + pixoff + 2) * 16);
      slide_bf<2, 2, 1>(
          w,
          [&](int i, uint32_t(&a)[2][4]) {
            ldsm4(a[0], aA + (uint32_t)(i * (P2 * 16)));
            ldsm2(a[1][0], a[1][1], aB + (uint32_t)(i * (P2 * 16)));
            a[1][2] = a[1][3] = 0u;
          },
          [&](int, int, const float(&top)[4], const float(&bot)[4]) {
            float sm4[4];
            sum2x2(top, bot, sm4);
            if (!(g & 1)) {
              const int py = warp - 4, px = g >> 1;
#pragma unroll
              for (int h = 0; h < 2; ++h)
                *reinterpret_cast<uint32_t*>(smraw + kDO2 + ((py + 1) * P3 + px + 4 * h + 1) * 16 + 4 * t) = pack_bf16(sm4[2 * h], sm4[2 * h + 1]);
            }
          });
    }
    __syncthreads();
    HG_MARK(24);
    // ================= D2: dec[2] (cat(e2, up(o3)) 24 -> 8 on 8x8): weight gradient, 15 (channel block, tap pair) tiles (warps 0-7)
    //                   || input gradient into up(o3) -> 2x2 sum -> d o3 (warps 8-15: 4 row pairs x 2 channel tiles)
    if (warp < 8) {
      for (int tile = warp; tile < 15; tile += 8) {
        const int cb = tile / 5, tp = tile - 5 * cb;
        const int tap = min(2 * tp + tsel, 8), ky = tap / 3, kx = tap - 3 * ky;
        float4* acc4 = reinterpret_cast<float4*>(sAcc + (T_D2 + tile) * 128) + lane;
        float4 cv = *acc4;
        float c[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const int vy = 2 * mt + thalf + ky, vx = lr + kx;            // haloed fine coordinates of the tapped pixel
          uint32_t a[4];
          ldsm4t(a, cb == 0 ? smb + tE2 + (uint32_t)((vy * P3 + vx) * 16)
                            : smb + tO3 + (uint32_t)((cb - 1) * PB4 + (((vy + 1) >> 1) * P4 + ((vx + 1) >> 1)) * 16));
          if (tp == 4) a[1] = a[3] = cb == 0 ? ones : 0u;
          uint32_t b0, b1;
          ldsm2t(b0, b1, smb + kDO2 + (uint32_t)(((2 * mt + (lj & 1) + 1) * P3 + 1 + lr) * 16));
          mma_bf16(c, a, b0, b1);
        }
        *acc4 = make_float4(c[0], c[1], c[2], c[3]);
      }
    } else {
      const int mt = (warp - 8) >> 1, nt = warp & 1;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int tp = 0; tp < 5; ++tp) {
        const int tap = min(2 * tp + chunk, 8), ky = tap / 3, kx = tap - 3 * ky;
        uint32_t a[4];
        ldsm4(a, smb + kDO2 + (uint32_t)(((2 * mt + (lj & 1) + ky) * P3 + lr + kx) * 16));
        const uint2 w = __ldg(p.pack + (B_D2D + tp * 2 + nt) * 32 + lane);
        mma_bf16(c, a, w.x, w.y);
      }
      // rows g: pixel (2mt, x = g); rows g+8: pixel (2mt+1, x = g): window = the two rows x the x-pair
      float s0 = c[0] + c[2], s1 = c[1] + c[3];
      s0 += __shfl_xor_sync(0xffffffffu, s0, 4);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 4);
      if (!(g & 1))
        *reinterpret_cast<uint32_t*>(smraw + kDO3 + nt * PB4 + ((mt + 1) * P4 + (g >> 1) + 1) * 16 + 4 * t) = pack_bf16(s0, s1);
    }
    __syncthreads();
    HG_MARK(25);
    // ================= D3: dec[3] (cat(e3, up4(dec[4])) 48 -> 16 on 4x4): weight gradient, 60 tiles of ONE MMA per frame;
    //                   input gradient into the dec[4] half, summed over the 16 pixels -> d(dec[4] output) (warps 0-3 afterwards)
    for (int tile = warp; tile < 60; tile += 16) {
      const int cb = tile / 10, tp = (tile % 10) >> 1, nt = tile & 1;
      const int tap = min(2 * tp + tsel, 8), ky = tap / 3, kx = tap - 3 * ky;
      float4* acc4 = reinterpret_cast<float4*>(sAcc + (T_D3 + tile) * 128) + lane;
      float4 cv = *acc4;
      float c[4] = {cv.x, cv.y, cv.z, cv.w};
      uint32_t a[4];
      ldsm4t(a, smb + tC3 + (uint32_t)(cb * PB4 + (((tpix >> 2) + ky) * P4 + (tpix & 3) + kx) * 16));
      if (tp == 4) a[1] = a[3] = cb == 0 ? ones : 0u;
      uint32_t b0, b1;
      const int bp = lr + 8 * (lj & 1);
      ldsm2t(b0, b1, smb + kDO3 + (uint32_t)(nt * PB4 + (((bp >> 2) + 1) * P4 + (bp & 3) + 1) * 16));
      mma_bf16(c, a, b0, b1);
      *acc4 = make_float4(c[0], c[1], c[2], c[3]);
    }
    if (warp < 4) {
      const int nt = warp;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const int ky = tp / 3, kx = tp - 3 * ky;
        uint32_t a[4];
        ldsm4(a, smb + kDO3 + (uint32_t)(chunk * PB4 + (((pixoff >> 2) + ky) * P4 + (pixoff & 3) + kx) * 16));
        const uint2 w = __ldg(p.pack + (B_D3D + tp * 4 + nt) * 32 + lane);
        mma_bf16(c, a, w.x, w.y);
      }
      float s0 = c[0] + c[2], s1 = c[1] + c[3];         // nearest x4 of a 1x1 map: the gradient is the sum over all 16 pixels
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
      if (g == 0) { sDD4[nt * 8 + 2 * t] = s0; sDD4[nt * 8 + 2 * t + 1] = s1; }
    }
    __syncthreads();
    HG_MARK(26);
    // ================= D4: dec[4] (1x1 conv on the bottleneck h): dW[o][i] += dD4[o] * h[i]
    accD4[0] = fmaf(sDD4[tid >> 5], sH[tid & 31], accD4[0]);
    accD4[1] = fmaf(sDD4[16 + (tid >> 5)], sH[tid & 31], accD4[1]);
    if (tid < 32) accB4 += sDD4[tid];
    HG_MARK(27);
    ++fr;
    if (p.dbg && n == 0) {                             // debug dump of frame 0: d o0 | d o1 | d o2 | d o3 planes (raw bytes) | dD4
      uint32_t* d = reinterpret_cast<uint32_t*>(p.dbg);
      const uint32_t* s = reinterpret_cast<const uint32_t*>(smraw + kDO0);
      for (int e = tid; e < (kW - kDO0) / 4; e += NT) d[e] = s[e];
      if (tid < 32) p.dbg[(kW - kDO0) / 4 + tid] = sDD4[tid];
    }
  }

  // ---- end of the CTA's frames: assemble the partial gradient vector (state_dict order) in the tape region.  Register
  //      accumulators go through scratch (the band region is dead); every destination has exactly ONE writer and copies of
  //      a K-split are summed in a fixed order: no atomics anywhere, the vector is bit-reproducible.
  __syncthreads();
  float* sG = reinterpret_cast<float*>(smraw);
  float* scrM0 = reinterpret_cast<float*>(smraw + kM0);            // [12 warps][3 ky][128]
  float* scrM2 = scrM0 + 12 * 3 * 128;                             // [16 warps][2 tiles][128] | [16] bias partials
  static_assert((12 * 3 * 128 + 16 * 2 * 128 + 16) * 4 <= 2 * PLB, "assembly scratch fits the band region");
  for (int e = tid; e < PSTRIDE_M; e += NT) sG[e] = 0.f;
  if (warp < 12) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
      *reinterpret_cast<float4*>(scrM0 + (warp * 3 + ky) * 128 + lane * 4) = make_float4(accM0[ky][0], accM0[ky][1], accM0[ky][2], accM0[ky][3]);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j)
    *reinterpret_cast<float4*>(scrM2 + (warp * 2 + j) * 128 + lane * 4) = make_float4(accM2[j][0], accM2[j][1], accM2[j][2], accM2[j][3]);
  bsum2 = warp_sum(bsum2);
  if (lane == 0) scrM2[16 * 2 * 128 + warp] = bsum2;
  __syncthreads();
  // fragment value q of lane (g, t): row m = g + 8*(q >> 1), column n = 2t + (q & 1)
  for (int e = tid; e < 6 * 3 * 128; e += NT) {          // masker.0 [16][11][3][3] + bias: role = nt*3 + tr, two row halves
    const int role = e / 384, ky = (e >> 7) % 3, ln = (e >> 2) & 31, q = e & 3, gg = ln >> 2, tt = ln & 3;
    const int tr = role % 3, nt = role / 3, m = gg + 8 * (q >> 1), co = nt * 8 + 2 * tt + (q & 1);
    const float v = scrM0[(role * 3 + ky) * 128 + ln * 4 + q] + scrM0[((role + 6) * 3 + ky) * 128 + ln * 4 + q];
    if (tr == 0) {
      const int kx = m >> 2, c = m & 3;
      if (kx < 3 && c < 3) sG[gM0W + ((co * 11 + c) * 3 + ky) * 3 + kx] = v;
    } else if (tr == 1) {
      sG[gM0W + ((co * 11 + 3 + (m & 7)) * 3 + ky) * 3 + (m >> 3)] = v;
    } else if (m < 8) {
      sG[gM0W + ((co * 11 + 3 + m) * 3 + ky) * 3 + 2] = v;
    } else if (m == 8 && ky == 0) {
      sG[gM0B + co] = v;
    }
  }
  if (tid < 256) {                                       // masker.2 [1][16][3][3]: row = ci, column = tap (tile 1: tap 8 in column 0)
    const int j = tid >> 7, ln = (tid >> 2) & 31, q = tid & 3, ci = (ln >> 2) + 8 * (q >> 1), col = 2 * (ln & 3) + (q & 1);
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) v += scrM2[(w * 2 + j) * 128 + ln * 4 + q];
    if (j == 0) sG[gM2W + ci * 9 + col] = v;
    else if (col == 0) sG[gM2W + ci * 9 + 8] = v;
  }
  if (tid == 0) {
    float v = 0.f;
    for (int w = 0; w < 16; ++w) v += scrM2[16 * 2 * 128 + w];
    sG[gM2B] = v;
  }
  sG[gD4W + tid] = accD4[0];                         // dec[4] [32][32] + bias
  sG[gD4W + 512 + tid] = accD4[1];
  if (tid < 32) sG[gD4B + tid] = accB4;
  // dec[0]: tiles [kh][tr][ky] (two row halves, summed here); dec[1]: tiles [tr][ky]; tr = src*2 + kxg
  for (int e = tid; e < 24 * 128; e += NT) {
    const int tile = e >> 7, ln = (e >> 2) & 31, q = e & 3, gg = ln >> 2, tt = ln & 3;
    const bool d0 = tile < 12;
    const int tl = d0 ? tile : tile - 12, tr = tl / 3, ky = tl % 3, src = tr >> 1, kxg = tr & 1;
    const int m = gg + 8 * (q >> 1), co = 2 * tt + (q & 1);
    const float v = d0 ? sAcc[(T_D0 + tile) * 128 + ln * 4 + q] + sAcc[(T_D0 + 12 + tile) * 128 + ln * 4 + q]
                       : sAcc[(T_D1 + tl) * 128 + ln * 4 + q];
    float* gw = sG + (d0 ? gD0W : gD1W);
    float* gb = sG + (d0 ? gD0B : gD1B);
    if (kxg == 0) gw[((co * 16 + src * 8 + (m & 7)) * 3 + ky) * 3 + (m >> 3)] = v;
    else if (m < 8) gw[((co * 16 + src * 8 + m) * 3 + ky) * 3 + 2] = v;
    else if (m == 8 && src == 0 && ky == 0) gb[co] = v;
  }
  // dec[2] [8][24][3][3]: tiles cb*5 + tp; rows m < 8: tap 2tp, m >= 8: tap 2tp+1 (tp 4: row 8 = bias on cb 0)
  for (int e = tid; e < 15 * 128; e += NT) {
    const int tile = e >> 7, ln = (e >> 2) & 31, q = e & 3, gg = ln >> 2, tt = ln & 3;
    const int cb = tile / 5, tp = tile - 5 * cb, m = gg + 8 * (q >> 1), co = 2 * tt + (q & 1), tap = 2 * tp + (m >> 3);
    const float v = sAcc[(T_D2 + tile) * 128 + ln * 4 + q];
    if (tap < 9) sG[gD2W + (co * 24 + cb * 8 + (m & 7)) * 9 + tap] = v;
    else if (m == 8 && cb == 0) sG[gD2B + co] = v;
  }
  // dec[3] [16][48][3][3]: tiles (cb*5 + tp)*2 + nt
  for (int e = tid; e < 60 * 128; e += NT) {
    const int tile = e >> 7, ln = (e >> 2) & 31, q = e & 3, gg = ln >> 2, tt = ln & 3;
    const int cb = tile / 10, tp = (tile % 10) >> 1, nt = tile & 1, m = gg + 8 * (q >> 1), co = nt * 8 + 2 * tt + (q & 1), tap = 2 * tp + (m >> 3);
    const float v = sAcc[(T_D3 + tile) * 128 + ln * 4 + q];
    if (tap < 9) sG[gD3W + (co * 48 + cb * 8 + (m & 7)) * 9 + tap] = v;
    else if (m == 8 && cb == 0) sG[gD3B + co] = v;
  }
  __syncthreads();
  {
    float4* d = reinterpret_cast<float4*>(p.partials + (size_t)blockIdx.x * PSTRIDE_M);
    const float4* s = reinterpret_cast<const float4*>(sG);
    for (int e = tid; e < PSTRIDE_M / 4; e += NT) d[e] = s[e];
  }
}

}  // namespace hg
}  // namespace cgs

using namespace cgs;

// Debug: point the phase traces of the forward / backward kernels at device buffers of 64 int64 each (NULL disables).
extern "C" int cgs_hg_set_trace(long long* fwd_buf, long long* bwd_buf, long long* score_buf) {
  if (hg::set_fwd_trace(fwd_buf) != 0 || hs::set_trace(score_buf) != 0) return -2;
  return cudaMemcpyToSymbol(hg::g_hgb_trace, &bwd_buf, sizeof(bwd_buf)) == cudaSuccess ? 0 : -2;
}

extern "C" int cgs_hg_grid(int32_t B) {
  if (B <= 0) return 0;
  const int sms = device_sms(), per = (B + sms - 1) / sms;
  return (B + per - 1) / per;
}
// Non-zero if a bounded mbarrier wait of the Hourglass kernels ever timed out (reads a device flag; synchronises).
extern "C" int cgs_hg_status(void) {
  int v = 0;
  if (cudaMemcpyFromSymbol(&v, hg::g_hg_timeout, sizeof(v)) != cudaSuccess) return -2;
  return v;
}

extern "C" int cgs_hg_debug_floats(void) { return (hg::kW - hg::kDO0) / 4 + 32; }

extern "C" int cgs_hg_backward(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const cgs_masker_weights* mw,
                               const uint32_t* pack, const void* tape, const float* z, const float* dz, float* partials,
                               float* debug, void* stream) {
  CGS_REQUIRE(frames && mw && mw->bm0 && pack && tape && z && dz && partials && B > 0, "hg_backward: bad args");
  CGS_REQUIRE((((uintptr_t)frames | (uintptr_t)pack | (uintptr_t)tape | (uintptr_t)partials) & 15) == 0,
              "hg_backward: frames, pack, tape and partials must be 16-byte aligned");
  hg::BwdParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames; p.tape = (const uint8_t*)tape; p.z = z; p.dz = dz; p.pack = reinterpret_cast<const uint2*>(pack);
  p.bm0 = mw->bm0; p.roll_dev = roll_dev; p.B = B; p.roll = roll; p.partials = partials; p.dbg = debug;
  cudaFuncSetAttribute(hg::hg_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hg::B_SMEM);
  hg::hg_backward_kernel<<<cgs_hg_grid(B), hg::NT, hg::B_SMEM, (cudaStream_t)stream>>>(p);
  return check_launch("hg_backward");
}
