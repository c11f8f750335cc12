// The wide (chfak > 1) convolution kernels: TMA-fed tcgen05 / TMEM implicit GEMMs with bf16 operands, fp32 accumulation.
//
// Activations live in HBM "chunk-planar": [B][C/8][H][W][8 channels] bf16, so that one pixel's 8 channels of a plane are one
// 16-byte slot.  A TMA box (8 bw, bh, planes, 1) of the 4-D view (8 W, H, C/8, B) lands in shared memory as
// [plane][bh][bw][16 B], which is at once
//   * the UMMA no-swizzle K-MAJOR canonical layout ((8,m),(8,2)):((16 B,SBO),(1,LBO)) with a core matrix = 8 horizontally
//     adjacent pixels of one plane, SBO = one tile row and LBO = one plane: the A operand (rows = pixels, K = channels) of the
//     forward / input-gradient GEMM.  The tile is loaded ONCE with its 1-pixel halo (TMA zero-fills out of bounds = the
//     convolution's padding) and filter tap (ky, kx) is the same buffer addressed (ky * 10 + kx) slots further on: nine shifted
//     descriptors, no im2col, no data movement between taps;
//   * the UMMA no-swizzle MN-MAJOR canonical layout ((8,m),(8,k)):((1,SBO),(16 B,LBO)) with SBO = one plane and LBO = one tile
//     row: both operands (K = pixels) of the weight-gradient GEMM  D[(kx, ci), co] = sum_pixels X[p + tap][ci] * dY[p][co].
//     Three TMA loads put the kx = 0, 1, 2 shifts of the input tile into planes [kx][ci / 8] of ONE 128-row A operand (plane 15
//     holds ones: the bias gradient rides along), so a tile costs 3 (ky) x 8 (row pairs) MMAs of 128 x Cout x 16.
//
// Both kernels are persistent and warp-specialised: warp 0 = TMA producer (one thread, `cp.async.bulk.tensor.4d` + mbarrier
// expect_tx, 3-4 stage ring), warp 1 = MMA issuer (one thread, `tcgen05.mma.cta_group::1.kind::f16`, `tcgen05.commit` frees
// the stage / publishes the accumulator), warps 2-5 = epilogue (tcgen05.ld of their TMEM lane quarter; bias + ReLU + 2x2
// max-pool + first-max arg-max + dropout mask, or the pool / ReLU / dropout BACKWARD scatter, written as bf16 planes).
// The accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the MMAs of tile i+1 and the loads of i+2...
// Reference: nets.py:169-187 (features), main.py:185-198 (the training step these feed).
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include "common.cuh"

namespace cgs {
namespace wd {

constexpr int TW = 8, TH = 16, HWID = 10, HHGT = 18;
constexpr int SLOTS = HWID * HHGT;                 // 180 haloed pixels
constexpr int NTHR = 320;                          // conv kernel: TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int NTHR_W = 192;                        // weight-gradient kernel: TMA warp, MMA warp, 4 warps that drain TMEM once at the end
constexpr int A3_ROWS = 18 * 8;                    // weight gradient: one kx copy of a plane = 18 rows x 8 pixels
constexpr int A3_PLANE = A3_ROWS * 16, A3_BYTES = 16 * A3_PLANE;
constexpr int DY_PLANE = TH * TW * 16;

__device__ int g_wd_timeout = 0;
__device__ long long* g_wd_trace = nullptr;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle shared-memory matrix descriptor (sm_100 format, version 1); the meaning of LBO / SBO depends on the major-ness
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must surface as an error flag (cgs_wide_status), never as a hung GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return true;
  }
  atomicExch(&g_wd_timeout, 1);
  return false;
}
// box (8 * bw, bh, planes, 1) of the 4-D view (8 * W, H, C/8, B): x in pixels, every box row one contiguous run of 16 * bw bytes
__device__ __forceinline__ void tma_load4(uint32_t dst, const CUtensorMap* tm, int x, int y, int plane, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dst),
      "l"(tm), "r"(8 * x), "r"(y), "r"(plane), "r"(n), "r"(bar)
      : "memory");
}
// n / d through one multiply-high (common.cuh FastDiv); d == 1 is the identity
__device__ __forceinline__ int udiv(int n, const FastDiv& f) { return f.d == 1 ? n : fdiv(n, f); }
static FastDiv fastdiv_or_one(int d) { return d >= 2 ? make_fastdiv(d) : FastDiv{0u, 1}; }

// one lane of a converged warp; unlike `lane == 0` the compiler knows the region is executed by a single lane, so the warp-uniform
// descriptor arithmetic of the MMA loop can stay in uniform registers (UTCHMMA takes its operands from there)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// one 16-byte row of the packed B operand: 8 input channels (ci = q * 8 ..) of output channel co at filter tap t
__device__ __forceinline__ uint4 pack_row(const float* __restrict__ w, int r, int KP, int NP, int Cin, int Cout, int transposed) {
  const int co = r % NP, tq = r / NP, q = tq % KP, t = tq / KP;
  float v[8];
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    const int ci = q * 8 + c8;
    v[c8] = 0.f;
    if (ci < Cin && co < Cout) v[c8] = transposed ? __ldg(w + ((size_t)ci * Cout + co) * 9 + (8 - t)) : __ldg(w + ((size_t)co * Cin + ci) * 9 + t);
  }
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

struct PackJob {
  const float* w;
  uint4* out;
  int Cin, Cout, transposed;       // GEMM view, as in cgs_wide_conv3x3
};
struct PackJobs {
  PackJob j[8];
};
// all operand tiles of a step in one launch (blockIdx.y = job): the filters change with every optimizer step
__global__ void wide_pack_kernel(const PackJobs jobs) {
  pdl_trigger();
  pdl_wait();
  const PackJob jb = jobs.j[blockIdx.y];
  const int KP = ((jb.Cin >> 3) + 1) & ~1, NP = (jb.Cout + 15) & ~15, rows = 9 * KP * NP;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) jb.out[r] = pack_row(jb.w, r, KP, NP, jb.Cin, jb.Cout, jb.transposed);
}

struct ConvP {
  int B, H, W, Cin, Cout;          // GEMM view: K = 9 taps x Cin, N = Cout (transposed: Cin = the layer's outputs)
  int CPi, KP, NP;                 // Cin / 8, planes padded to even, Cout padded to 16
  int tiles_x, tiles_y, ntiles, stages, tmem_cols, epi, transposed;
  FastDiv d_tpf, d_tx;             // dividers by tiles per frame / tiles per row
  uint32_t idesc;
  const float* w;
  const uint4* wpacked;            // [9][KP][NP] x 16 B operand rows (wide_pack_kernel), or NULL: packed here from w
  const float* bias;
  __nv_bfloat16* out;
  uint8_t* idx_out;
  const uint8_t* idx_in;
  const float* mask;
  float* out_f32;
};

template <int KS>
__global__ void __launch_bounds__(NTHR, 1) wide_conv_kernel(const __grid_constant__ CUtensorMap tmx, const ConvP p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KP = p.KP, NP = p.NP;
  pdl_trigger();
  if (blockIdx.x == 0 && tid == 0 && g_wd_trace) g_wd_trace[56] = clock64();
  const uint32_t w_bytes = (uint32_t)9 * KP * NP * 16, a_bytes = (uint32_t)KP * SLOTS * 16, a_tx = (uint32_t)p.CPi * SLOTS * 16;
  unsigned char* s_w = smem;
  unsigned char* s_a = smem + w_bytes;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_a + (size_t)p.stages * a_bytes);     // full[S] empty[S] tfull[2] tempty[2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * p.stages + 4);
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t bFull = bar0, bEmpty = bar0 + 8 * p.stages, bTFull = bar0 + 16 * p.stages, bTEmpty = bTFull + 16;

  // ---- one-time setup
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bFull + 8 * s, 1); mbar_init(bEmpty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bTFull + 8 * b, 1); mbar_init(bTEmpty + 8 * b, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  pdl_wait();                                      // from here on: global memory the kernels before this one wrote
  {
    // B operand (K-major): element (tap, ci, co) -> [(tap * KP + ci / 8) * NP + co] x 16 B + (ci % 8) x 2 B, zero padded
    const int rows = 9 * KP * NP;
    if (p.wpacked) {
      for (int r = tid; r < rows; r += NTHR) reinterpret_cast<uint4*>(s_w)[r] = __ldg(p.wpacked + r);
    } else {
      for (int r = tid; r < rows; r += NTHR) reinterpret_cast<uint4*>(s_w)[r] = pack_row(p.w, r, KP, NP, p.Cin, p.Cout, p.transposed);
    }
    // the pad plane of every stage (Cin / 8 odd) is read by the last MMA of a tap: its weights are zero, the plane must be finite
    if (KP > p.CPi)
      for (int s = 0; s < p.stages; ++s)
        for (int e = tid; e < SLOTS; e += NTHR) reinterpret_cast<uint4*>(s_a + (size_t)s * a_bytes + (size_t)p.CPi * SLOTS * 16)[e] = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const int tpf = p.tiles_x * p.tiles_y;
  long long* trace = (blockIdx.x == 0) ? g_wd_trace : nullptr;
  if (trace && tid == 0) trace[57] = clock64();

  if (warp == 0) {
    // ===== TMA producer
    if (elect_one()) {
      int s = 0, k = 0;                             // stage, ring pass (no divisions on these single-thread paths)
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        if (!mbar_wait(bEmpty + 8 * s, (k & 1) ^ 1)) break;
        const int n = udiv(tile, p.d_tpf), r = tile - n * tpf, ty = udiv(r, p.d_tx), tx = r - ty * p.tiles_x;
        mbar_expect_tx(bFull + 8 * s, a_tx);
        tma_load4(smem_u32(s_a + (size_t)s * a_bytes), &tmx, tx * TW - 1, ty * TH - 1, 0, n, bFull + 8 * s);
        if (++s == p.stages) { s = 0; ++k; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (elect_one()) {
      const uint64_t a_desc0 = umma_desc(smem_u32(s_a), SLOTS * 16, HWID * 16), b_desc0 = umma_desc(smem_u32(s_w), (uint32_t)NP * 16, 128);
      const uint64_t b_step = (uint64_t)(2 * NP);   // consecutive (tap, plane pair) operand tiles are 2 NP rows of 16 bytes apart
      int it = 0, s = 0, k = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int ab = it & 1, j = it >> 1;
        if (trace && it < 8) trace[it * 8 + 0] = clock64();
        if (!mbar_wait(bTEmpty + 8 * ab, (j & 1) ^ 1)) break;
        if (trace && it < 8) trace[it * 8 + 1] = clock64();
        if (!mbar_wait(bFull + 8 * s, k & 1)) break;
        if (trace && it < 8) trace[it * 8 + 2] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // the 9 * KS descriptors of a tile differ in their start-address fields only: one 64-bit add each, everything else is
        // loop-invariant (a single thread issues these: its own instruction latency is what bounds the MMA rate)
        const uint64_t ad0 = a_desc0 + (uint64_t)((uint32_t)s * (a_bytes >> 4));
        const uint32_t td = tmem_base + (uint32_t)(ab * NP);
        uint64_t bd = b_desc0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            umma_bf16(td, ad0 + (uint64_t)((tap / 3) * HWID + (tap % 3) + 2 * ks * SLOTS), bd, p.idesc, (tap | ks) ? 1u : 0u);
            bd += b_step;
          }
        }
        umma_commit(bEmpty + 8 * s);              // the stage is free once these MMAs have read it
        umma_commit(bTFull + 8 * ab);             // ... and the accumulator is complete
        if (trace && it < 8) trace[it * 8 + 3] = clock64();
        if (++s == p.stages) { s = 0; ++k; }
      }
      if (trace) trace[58] = clock64();
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter (warp & 3); lane m of the accumulator = pixel (m / 8, m % 8) of the tile
    const int q = warp & 3, half = (warp - 2) >> 2, m = q * 32 + lane, yl = m >> 3, xl = m & 7;
    const bool odd_x = xl & 1, odd_y = yl & 1;
    const int H = p.H, W = p.W, Cout = p.Cout, CPo = Cout >> 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int ab = it & 1, j = it >> 1;
      if (trace && it < 8 && tid == 64) trace[it * 8 + 4] = clock64();
      if (!mbar_wait(bTFull + 8 * ab, j & 1)) break;
      if (trace && it < 8 && tid == 64) trace[it * 8 + 5] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int n = udiv(tile, p.d_tpf), r = tile - n * tpf, ty = udiv(r, p.d_tx), tx = r - ty * p.tiles_x;
      const int y = ty * TH + yl, x = tx * TW + xl;
      const bool inb = y < H;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * NP);
      for (int cg = half; cg < CPo; cg += 2) {
        float a[8];
        tmem_ld8(taddr + (uint32_t)(cg * 8), a);
        if (p.bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8) + 1);
          a[0] += b0.x; a[1] += b0.y; a[2] += b0.z; a[3] += b0.w; a[4] += b1.x; a[5] += b1.y; a[6] += b1.z; a[7] += b1.w;
        }
        if (p.epi == CGS_WIDE_EPI_RELU_POOL) {
          // 2x2 window = lanes (l, l^1, l^8, l^9); the first maximum in row-major order wins (ATen max_pool2d); 4 = ReLU is dead.
          // The four lanes of a window split its 8 channels: after the x exchange a lane holds 4 channels (even x: 0-3, odd x:
          // 4-7), after the y exchange 2 (even y: the first two of those): 7 shuffles per 8 channels, every lane stores.
          float mx[4];
          uint32_t rb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float recv = __shfl_xor_sync(0xffffffffu, odd_x ? a[i] : a[4 + i], 1), mine = odd_x ? a[4 + i] : a[i];
            const float left = odd_x ? recv : mine, right = odd_x ? mine : recv;
            rb[i] = right > left ? 1u : 0u;
            mx[i] = rb[i] ? right : left;
          }
          const uint32_t mybits = odd_y ? (rb[2] | (rb[3] << 1)) : (rb[0] | (rb[1] << 1));
          const uint32_t obits = __shfl_xor_sync(0xffffffffu, odd_y ? (rb[0] | (rb[1] << 1)) : (rb[2] | (rb[3] << 1)), 8);
          const uint32_t tbits = odd_y ? obits : mybits, bbits = odd_y ? mybits : obits;
          float res[2];
          uint32_t am[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const float recv = __shfl_xor_sync(0xffffffffu, odd_y ? mx[k] : mx[2 + k], 8), mine = odd_y ? mx[2 + k] : mx[k];
            const float top = odd_y ? recv : mine, bot = odd_y ? mine : recv;
            const bool bb = bot > top;
            res[k] = bb ? bot : top;
            am[k] = bb ? (2u + ((bbits >> k) & 1u)) : ((tbits >> k) & 1u);
            if (!(res[k] > 0.f)) { res[k] = 0.f; am[k] = 4u; }
          }
          if (inb) {
            const int H2 = H >> 1, W2 = W >> 1, y2 = y >> 1, x2 = x >> 1, c0 = (odd_x ? 4 : 0) + (odd_y ? 2 : 0);
            if (p.mask) {
              const float2 mk = __ldg(reinterpret_cast<const float2*>(p.mask + ((size_t)(n * H2 + y2) * W2 + x2) * Cout + cg * 8 + c0));
              res[0] *= mk.x; res[1] *= mk.y;
            }
            const size_t o = (((size_t)(n * CPo + cg) * H2 + y2) * W2 + x2) * 8 + c0;
            if (p.out) *reinterpret_cast<uint32_t*>(p.out + o) = pack2(res[0], res[1]);
            if (p.out_f32) {       // NCHW fp32 copy for the head: features.14 reads a frame's [C][4][4] block as its K vector
              p.out_f32[((size_t)(n * Cout + cg * 8 + c0) * H2 + y2) * W2 + x2] = res[0];
              p.out_f32[((size_t)(n * Cout + cg * 8 + c0 + 1) * H2 + y2) * W2 + x2] = res[1];
            }
            *reinterpret_cast<unsigned short*>(p.idx_out + o) = (unsigned short)(am[0] | (am[1] << 8));
          }
        } else if (p.epi == CGS_WIDE_EPI_UNPOOL) {
          // this pixel is one pooled output of the layer below: its gradient goes to the arg-max position of the 2x2 window
          // (ReLU dead / dropout: nowhere / scaled), the other three positions get zeros
          if (inb) {
            const size_t oi = ((size_t)(n * CPo + cg) * H + y) * W + x;
            const uint2 ib = __ldg(reinterpret_cast<const uint2*>(p.idx_in) + oi);
            if (p.mask) {
              const float* mp = p.mask + ((size_t)(n * H + y) * W + x) * Cout + cg * 8;
              const float4 m0 = __ldg(reinterpret_cast<const float4*>(mp)), m1 = __ldg(reinterpret_cast<const float4*>(mp) + 1);
              a[0] *= m0.x; a[1] *= m0.y; a[2] *= m0.z; a[3] *= m0.w; a[4] *= m1.x; a[5] *= m1.y; a[6] *= m1.z; a[7] *= m1.w;
            }
            const int H2 = 2 * H, W2 = 2 * W;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              uint32_t pk[2][4];
#pragma unroll
              for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) {
                  const uint32_t i0 = ((2 * c2 < 4 ? ib.x : ib.y) >> (8 * ((2 * c2) & 3))) & 0xffu;
                  const uint32_t i1 = ((2 * c2 + 1 < 4 ? ib.x : ib.y) >> (8 * ((2 * c2 + 1) & 3))) & 0xffu;
                  const uint32_t pos = (uint32_t)(dy * 2 + dx);
                  pk[dx][c2] = pack2(i0 == pos ? a[2 * c2] : 0.f, i1 == pos ? a[2 * c2 + 1] : 0.f);
                }
              uint4* d = reinterpret_cast<uint4*>(p.out) + ((size_t)(n * CPo + cg) * H2 + 2 * y + dy) * W2 + 2 * x;
              d[0] = make_uint4(pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
              d[1] = make_uint4(pk[1][0], pk[1][1], pk[1][2], pk[1][3]);
            }
          }
        } else {
          if (inb) {
            const size_t o = ((size_t)(n * CPo + cg) * H + y) * W + x;
            reinterpret_cast<uint4*>(p.out)[o] = make_uint4(pack2(a[0], a[1]), pack2(a[2], a[3]), pack2(a[4], a[5]), pack2(a[6], a[7]));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bTEmpty + 8 * ab);
      if (trace && it < 8 && tid == 64) trace[it * 8 + 6] = clock64();
    }
    if (trace && tid == 64) trace[59] = clock64();
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (trace && tid == 0) trace[60] = clock64();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// --------------------------------------------------------------------------------------------------------------------
// weight gradient: D_ky[(kx, ci) | ones][co] += sum over the tile's pixels; accumulators stay in TMEM over all tiles of the CTA
struct WgP {
  int B, H, W, Cin, Cout, CPi, CPo, NP, NPl;
  int tiles_x, tiles_y, ntiles, stages, tmem_cols;
  FastDiv d_tpf, d_tx;
  uint32_t idesc;
  float* partials;                 // [grid][3][128][NP]
};

__global__ void __launch_bounds__(NTHR_W, 1) wide_wgrad_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmdy,
                                                             const WgP p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NP = p.NP;
  const uint32_t b_bytes = (uint32_t)p.NPl * DY_PLANE, st_bytes = A3_BYTES + b_bytes;
  const uint32_t tx_bytes = (uint32_t)3 * p.CPi * A3_PLANE + (uint32_t)p.CPo * DY_PLANE;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * st_bytes);   // full[S] empty[S] done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * p.stages + 1);
  const uint32_t bar0 = smem_u32(s_bar), bFull = bar0, bEmpty = bar0 + 8 * p.stages, bDone = bar0 + 16 * p.stages;
  pdl_trigger();

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bFull + 8 * s, 1); mbar_init(bEmpty + 8 * s, 1); }
    mbar_init(bDone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // planes the TMA never writes: zeros (unused rows / columns of D stay finite), plane 15 of A = ones (bias gradient)
  for (int s = 0; s < p.stages; ++s) {
    uint4* a = reinterpret_cast<uint4*>(smem + (size_t)s * st_bytes);
    for (int e = tid; e < 16 * A3_ROWS; e += NTHR_W) {
      const int pl = e / A3_ROWS;
      if (pl >= 3 * p.CPi) a[e] = pl == 15 ? make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u) : make_uint4(0u, 0u, 0u, 0u);
    }
    uint4* b = reinterpret_cast<uint4*>(smem + (size_t)s * st_bytes + A3_BYTES);
    for (int e = tid + p.CPo * (TH * TW); e < p.NPl * (TH * TW); e += NTHR_W) b[e] = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const int tpf = p.tiles_x * p.tiles_y;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      int s = 0, k = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        if (!mbar_wait(bEmpty + 8 * s, (k & 1) ^ 1)) break;
        const int n = udiv(tile, p.d_tpf), r = tile - n * tpf, ty = udiv(r, p.d_tx), tx = r - ty * p.tiles_x;
        const uint32_t a0 = smem_u32(smem + (size_t)s * st_bytes);
        mbar_expect_tx(bFull + 8 * s, tx_bytes);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          tma_load4(a0 + (uint32_t)(kx * p.CPi) * A3_PLANE, &tmx, tx * TW - 1 + kx, ty * TH - 1, 0, n, bFull + 8 * s);
        tma_load4(a0 + A3_BYTES, &tmdy, tx * TW, ty * TH, 0, n, bFull + 8 * s);
        if (++s == p.stages) { s = 0; ++k; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // A: (LBO = one tile row, SBO = one plane of a kx copy); B: the dY tile behind it, same LBO, SBO = one dY plane.  The SBO
      // field is the only difference between the two descriptors besides the start address.
      const uint64_t a_desc0 = umma_desc(smem_u32(smem), 128, A3_PLANE);
      const uint64_t b_off = (umma_desc(smem_u32(smem) + A3_BYTES, 128, DY_PLANE) - a_desc0);
      int it = 0;
      bool ok = true;
      int s = 0, k = 0;
      for (int tile = blockIdx.x; tile < p.ntiles && ok; tile += gridDim.x, ++it) {
        if (!mbar_wait(bFull + 8 * s, k & 1)) { ok = false; break; }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint64_t ad0 = a_desc0 + (uint64_t)((uint32_t)s * (st_bytes >> 4)), bd0 = ad0 + b_off;
        const uint32_t first = it > 0 ? 1u : 0u;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int r2 = 0; r2 < 8; ++r2)          // K = 16 pixels = tile rows 2 r2, 2 r2 + 1 (8 pixels each): constant descriptor offsets
            umma_bf16(tmem_base + (uint32_t)(ky * NP), ad0 + (uint64_t)((2 * r2 + ky) * 8), bd0 + (uint64_t)(2 * r2 * 8), p.idesc, r2 ? 1u : first);
        umma_commit(bEmpty + 8 * s);
        if (++s == p.stages) { s = 0; ++k; }
      }
      umma_commit(bDone);
    }
  } else {
    const int q = warp & 3, m = q * 32 + lane;
    if (mbar_wait(bDone, 0)) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int ky = 0; ky < 3; ++ky) {
        float* dst = p.partials + (((size_t)blockIdx.x * 3 + ky) * 128 + m) * NP;
        for (int c = 0; c < NP; c += 8) {
          float a[8];
          tmem_ld8(taddr + (uint32_t)(ky * NP + c), a);
          reinterpret_cast<float4*>(dst + c)[0] = make_float4(a[0], a[1], a[2], a[3]);
          reinterpret_cast<float4*>(dst + c)[1] = make_float4(a[4], a[5], a[6], a[7]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// dW[co][ci][ky][kx] += sum over CTAs of P[cta][ky][kx * Cin + ci][co]; db[co] += sum of P[cta][1][120][co] (the ones plane).
// One warp per 4 consecutive elements: the lanes split the CTAs (independent 16-byte loads), then a fixed butterfly: deterministic.
__global__ void wide_wgrad_reduce_kernel(const float* __restrict__ part, int ncta, int NP, int Cin, int Cout, float* __restrict__ dw,
                                         float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  const int e4 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total4 = 3 * 128 * NP / 4;
  if (e4 >= total4) return;
  const int e = e4 * 4, co0 = e % NP, m = (e / NP) & 127, ky = e / (NP * 128);
  const bool is_w = m < 3 * Cin, is_b = (m == 120 && ky == 1 && db != nullptr);
  if ((!is_w && !is_b) || co0 >= Cout) return;                 // whole warp: same e4
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* p4 = reinterpret_cast<const float4*>(part) + e4;
  for (int c = lane; c < ncta; c += 32) {
    const float4 v = __ldg(p4 + (size_t)c * total4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
  }
  if (lane < 4) {
    const int co = co0 + lane;
    const float v = lane == 0 ? s.x : lane == 1 ? s.y : lane == 2 ? s.z : s.w;
    if (co < Cout) {
      if (is_w) {
        const int kx = m / Cin, ci = m - kx * Cin;
        dw[((size_t)(co * Cin + ci) * 3 + ky) * 3 + kx] += v;
      } else {
        db[co] += v;
      }
    }
  }
}

// ---- host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode_fn() {
  static EncodeFn f = nullptr;
  if (!f) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      f = (EncodeFn)ptr;
  }
  return f;
}

// 4-D view (8 * W, H, C/8, B) of a chunk-planar bf16 tensor; box (8 * bw, bh, planes, 1); out-of-bounds elements read as zero
int make_tmap(CUtensorMap* tm, const void* base, int B, int CP, int H, int W, int bw, int bh, int planes) {
  EncodeFn enc = encode_fn();
  if (!enc) {
    set_error("wide: cuTensorMapEncodeTiled is not available from this driver");
    return CGS_ECUDA;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)CP, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)CP * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)bw * 8, (cuuint32_t)bh, (cuuint32_t)planes, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("wide: cuTensorMapEncodeTiled failed (%d) for B=%d CP=%d H=%d W=%d box=(%d,%d,%d)", (int)r, B, CP, H, W, bw, bh, planes);
    return CGS_ECUDA;
  }
  return 0;
}

static int pow2_cols(int c) {
  int t = 32;
  while (t < c) t *= 2;
  return t;
}

int launch_wide_conv(const void* x, int B, int H, int W, int Cin, const float* w, const void* wpacked, const float* bias, int Cout, int transposed, int epi,
                     void* out, float* out_f32, uint8_t* idx_out, const uint8_t* idx_in, const float* mask, cudaStream_t st) {
  CGS_REQUIRE(x && (w || wpacked) && (out || out_f32) && B > 0, "wide_conv3x3: bad args");
  CGS_REQUIRE(((uintptr_t)wpacked & 15) == 0, "wide_conv3x3: packed weights must be 16-byte aligned");
  CGS_REQUIRE((Cin % 8) == 0 && (Cout % 8) == 0 && Cin >= 8 && Cout >= 8 && Cin <= 160 && Cout <= 240, "wide_conv3x3: channels must be multiples of 8 (Cin %d, Cout %d)", Cin, Cout);
  CGS_REQUIRE((W % TW) == 0 && (H % 8) == 0 && H >= 8, "wide_conv3x3: H x W = %d x %d unsupported", H, W);
  CGS_REQUIRE(epi != CGS_WIDE_EPI_RELU_POOL || idx_out, "wide_conv3x3: the pooling epilogue needs idx_out");
  CGS_REQUIRE(epi != CGS_WIDE_EPI_UNPOOL || (idx_in && out), "wide_conv3x3: the unpooling epilogue needs idx_in");
  ConvP p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.CPi = Cin / 8; p.KP = (p.CPi + 1) & ~1; p.NP = (Cout + 15) & ~15;
  p.tiles_x = W / TW; p.tiles_y = (H + TH - 1) / TH; p.ntiles = B * p.tiles_x * p.tiles_y;
  p.d_tpf = fastdiv_or_one(p.tiles_x * p.tiles_y); p.d_tx = fastdiv_or_one(p.tiles_x);
  p.epi = epi; p.transposed = transposed; p.w = w; p.wpacked = (const uint4*)wpacked; p.bias = bias;
  p.out = (__nv_bfloat16*)out; p.out_f32 = out_f32; p.idx_out = idx_out; p.idx_in = idx_in; p.mask = mask;
  p.tmem_cols = pow2_cols(2 * p.NP);
  CGS_REQUIRE(p.tmem_cols <= 512, "wide_conv3x3: Cout %d needs more than 512 TMEM columns", Cout);
  // D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t w_bytes = (size_t)9 * p.KP * p.NP * 16, a_bytes = (size_t)p.KP * SLOTS * 16;
  int stages = 4;
  while (stages > 2 && w_bytes + stages * a_bytes + 256 > 220 * 1024) --stages;
  p.stages = stages;
  const size_t smem = w_bytes + stages * a_bytes + 256;
  CGS_REQUIRE(smem <= 225 * 1024, "wide_conv3x3: Cin %d x Cout %d does not fit in shared memory", Cin, Cout);
  CUtensorMap tm;
  const int rc = make_tmap(&tm, x, B, p.CPi, H, W, HWID, HHGT, p.CPi);
  if (rc) return rc;
  int grid = device_sms();
  if (grid > p.ntiles) grid = p.ntiles;
#define CGS_WIDE_CONV(KS)                                                                                        \
  case KS:                                                                                                       \
    cudaFuncSetAttribute(wide_conv_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);         \
    launch_pdl(wide_conv_kernel<KS>, dim3(grid), dim3(NTHR), smem, st, tm, p);                                   \
    break;
  switch (p.KP / 2) {
    CGS_WIDE_CONV(1) CGS_WIDE_CONV(2) CGS_WIDE_CONV(3) CGS_WIDE_CONV(4) CGS_WIDE_CONV(5) CGS_WIDE_CONV(6) CGS_WIDE_CONV(7) CGS_WIDE_CONV(8)
    CGS_WIDE_CONV(9) CGS_WIDE_CONV(10)
    default:
      CGS_REQUIRE(false, "wide_conv3x3: Cin %d unsupported", Cin);
  }
#undef CGS_WIDE_CONV
  return check_launch("wide_conv3x3");
}

int wide_wgrad_grid(int B, int H, int W) {
  // at least ~4 tiles per CTA: every CTA hands over a 3 x 128 x NP partial tile, which is what the small layers would be paying for
  const int ntiles = B * (W / TW) * ((H + TH - 1) / TH);
  int grid = device_sms();
  const int want = (ntiles + 3) / 4;
  if (grid > want) grid = want;
  return grid < 1 ? 1 : grid;
}

int launch_wide_wgrad(const void* x, const void* dy, int B, int H, int W, int Cin, int Cout, float* dw, float* db, float* ws,
                      long long ws_floats, cudaStream_t st) {
  CGS_REQUIRE(x && dy && dw && ws && B > 0, "wide_wgrad3x3: bad args");
  CGS_REQUIRE((Cin % 8) == 0 && Cin >= 8 && Cin <= 40 && (Cout % 8) == 0 && Cout >= 8 && 3 * ((Cout + 15) & ~15) <= 512,
              "wide_wgrad3x3: Cin %d (<= 40) / Cout %d unsupported", Cin, Cout);
  CGS_REQUIRE((W % TW) == 0 && (H % 8) == 0 && H >= 8, "wide_wgrad3x3: H x W = %d x %d unsupported", H, W);
  WgP p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.CPi = Cin / 8; p.CPo = Cout / 8;
  p.NP = (Cout + 15) & ~15; p.NPl = p.NP / 8;
  p.tiles_x = W / TW; p.tiles_y = (H + TH - 1) / TH; p.ntiles = B * p.tiles_x * p.tiles_y;
  p.d_tpf = fastdiv_or_one(p.tiles_x * p.tiles_y); p.d_tx = fastdiv_or_one(p.tiles_x);
  p.tmem_cols = pow2_cols(3 * p.NP);
  // as above with A and B MN-major (bits 15, 16): K = pixels
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t st_bytes = (size_t)A3_BYTES + (size_t)p.NPl * DY_PLANE;
  int stages = 4;
  while (stages > 2 && stages * st_bytes + 256 > 220 * 1024) --stages;
  p.stages = stages;
  const size_t smem = stages * st_bytes + 256;
  const int grid = wide_wgrad_grid(B, H, W);
  CGS_REQUIRE(ws_floats >= (long long)grid * 3 * 128 * p.NP, "wide_wgrad3x3: workspace too small (%lld floats, need %lld)", ws_floats,
              (long long)grid * 3 * 128 * p.NP);
  p.partials = ws;
  CUtensorMap tmx, tmdy;
  int rc = make_tmap(&tmx, x, B, p.CPi, H, W, TW, HHGT, p.CPi);
  if (rc) return rc;
  rc = make_tmap(&tmdy, dy, B, p.CPo, H, W, TW, TH, p.CPo);
  if (rc) return rc;
  cudaFuncSetAttribute(wide_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  launch_pdl(wide_wgrad_kernel, dim3(grid), dim3(NTHR_W), smem, st, tmx, tmdy, p);
  rc = check_launch("wide_wgrad3x3");
  if (rc) return rc;
  const int nwarps = 3 * 128 * p.NP / 4;
  launch_pdl(wide_wgrad_reduce_kernel, dim3((nwarps + 7) / 8), dim3(256), 0, st, (const float*)ws, grid, p.NP, Cin, Cout, dw, db);
  return check_launch("wide_wgrad3x3.reduce");
}

}  // namespace wd
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_wide_conv3x3(const void* x, int32_t B, int32_t H, int32_t W, int32_t Cin, const float* w, const void* wpacked, const float* bias,
                                int32_t Cout, int32_t transposed, int32_t epi, void* out, float* out_f32, uint8_t* idx_out, const uint8_t* idx_in,
                                const float* mask, void* stream) {
  return wd::launch_wide_conv(x, B, H, W, Cin, w, wpacked, bias, Cout, transposed, epi, out, out_f32, idx_out, idx_in, mask, (cudaStream_t)stream);
}

extern "C" int64_t cgs_wide_packed_bytes(int32_t Cin, int32_t Cout) {
  return (int64_t)9 * (((Cin >> 3) + 1) & ~1) * ((Cout + 15) & ~15) * 16;
}

extern "C" int cgs_wide_pack(const cgs_wide_packjob* jobs, int32_t njobs, void* stream) {
  CGS_REQUIRE(jobs && njobs >= 1 && njobs <= 8, "wide_pack: 1..8 jobs");
  wd::PackJobs js;
  for (int i = 0; i < njobs; ++i) {
    CGS_REQUIRE(jobs[i].w && jobs[i].out && (jobs[i].Cin % 8) == 0 && (jobs[i].Cout % 8) == 0 && ((uintptr_t)jobs[i].out & 15) == 0, "wide_pack: bad job %d", i);
    js.j[i] = wd::PackJob{jobs[i].w, (uint4*)jobs[i].out, jobs[i].Cin, jobs[i].Cout, jobs[i].transposed};
  }
  launch_pdl(wd::wide_pack_kernel, dim3(16, njobs), dim3(256), 0, (cudaStream_t)stream, js);
  return check_launch("wide_pack");
}

extern "C" int cgs_wide_wgrad3x3(const void* x, const void* dy, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, float* dw, float* db,
                                 float* workspace, int64_t workspace_floats, void* stream) {
  return wd::launch_wide_wgrad(x, dy, B, H, W, Cin, Cout, dw, db, workspace, workspace_floats, (cudaStream_t)stream);
}

extern "C" int64_t cgs_wide_wgrad_workspace(int32_t B, int32_t H, int32_t W, int32_t Cout) {
  return (int64_t)wd::wide_wgrad_grid(B, H, W) * 3 * 128 * ((Cout + 15) & ~15);
}

extern "C" int cgs_wide_status(void) {
  int h = 0;
  cudaMemcpyFromSymbol(&h, wd::g_wd_timeout, sizeof(int));
  return h;
}

extern "C" int cgs_wide_set_trace(long long* dev_buf) {
  return cudaMemcpyToSymbol(wd::g_wd_trace, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -2;
}
