// tcgen05 / TMEM implicit-GEMM 3x3 convolution (fprop and dgrad), TF32 operands, fp32 accumulate.
//
// GEMM view per output tile:  D[128 pixels, Cout] = sum_{tap, ci} A_tap[128 pixels, ci] * W_tap[ci, Cout].
// The A operand is never materialised as im2col.  A tile of 16 x 8 output pixels is staged ONCE into
// shared memory with a 1-pixel halo (18 x 10 pixels), "chunk-planar": plane q holds channels [4q, 4q+4)
// of every haloed pixel as one 16-byte slot, slots row-major over the haloed tile.  In the UMMA
// no-swizzle K-major canonical layout ((8,m),(T,2)):((16B,SBO),(1,LBO)) a core matrix is 8 rows x 16 B
// = 8 horizontally adjacent pixels of one plane, SBO = one haloed row (160 B) and LBO = one plane, so
// the operand of filter tap (ky,kx) is just the SAME buffer addressed with the start address moved by
// (ky*10 + kx) slots: 9 shifted descriptors, zero data movement between taps, zero padding comes from
// the halo.  One elected thread issues the 9 * Cin/8 tcgen05.mma (M=128, N=Cout padded to 16, K=8) into
// a TMEM accumulator; tcgen05.commit arrives on an mbarrier; the four warps then read their 32 TMEM
// lanes (one output pixel per thread, all channels in registers) with tcgen05.ld and run the fused
// epilogue: bias, ReLU + 2x2 max-pool + argmax via lane shuffles (a pool window lives in lanes l, l^1,
// l^8, l^9), LeakyReLU, sigmoid + threshold, dropout-mask multiply, or the concat/upsample backward
// split.  Weights are repacked once per CTA into the matching B layout and stay resident; the CTA is
// persistent over tiles.  Operand prologues (concat+upsample, dropout, pool/ReLU/sigmoid/leaky
// backward, uint8 cast + roll) are the same src_load8 loaders as the fp32 kernel.
#include "common.cuh"

namespace cgs {

constexpr int TC_TW = 8, TC_TH = 16, TC_HW = 10, TC_HH = 18, TC_SLOTS = TC_HW * TC_HH;   // 180 haloed pixels
constexpr int TC_PLANE = TC_SLOTS * 16;                                                   // bytes per 4-channel plane
constexpr int TC_THREADS = 128;

struct TcGeom {
  int tiles_x, tiles_y, ntiles;
  int cin_pad, nch, n_pad, tmem_cols;
  uint32_t idesc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle shared-memory matrix descriptor (sm_100 format, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded spin: a protocol bug must surface as an error, never as a hung GPU
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

__global__ void __launch_bounds__(TC_THREADS) conv3x3_tc_kernel(const cgs_conv3x3_args p, const TcGeom g, int* __restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W, Cin = p.src.C, Cout = p.Cout;
  const int nch = g.nch, n_pad = g.n_pad;
  float* s_w = reinterpret_cast<float*>(smem_raw);                                   // [9][nch][n_pad][4]
  float* s_a = s_w + (size_t)9 * nch * n_pad * 4;                                    // [nch][180][4]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_a + (size_t)nch * TC_SLOTS * 4);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);

  // ---- one-time setup: TMEM allocation (warp 0), mbarrier, resident weights
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  {
    // B operand: element (tap, ci, co) -> [(tap*nch + ci/4)*n_pad + co]*4 + ci%4, zero padded
    const int total = 9 * nch * n_pad * 4;
    for (int e = tid; e < total; e += TC_THREADS) {
      const int c4 = e & 3, co = (e >> 2) % n_pad, r = (e >> 2) / n_pad;
      const int q = r % nch, t = r / nch;
      const int ci = q * 4 + c4;
      float v = 0.f;
      if (ci < Cin && co < Cout)
        v = p.transposed ? __ldg(p.w + ((size_t)ci * Cout + co) * 9 + (8 - t)) : __ldg(p.w + ((size_t)co * Cin + ci) * 9 + t);
      s_w[e] = v;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w), bar = smem_u32(s_bar);
  const uint32_t w_plane = (uint32_t)n_pad * 16;     // bytes between consecutive 4-channel weight planes
  const int ngroups = g.cin_pad >> 3;
  uint32_t phase = 0;

  // this thread's output pixel inside a tile: TMEM lane m = 32*warp + lane -> (yl, xl) = (m / 8, m % 8)
  const int m = tid, yl = m >> 3, xl = m & 7;
  const bool pool_owner = ((yl | xl) & 1) == 0;

  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    int t = tile;
    const int tix = t % g.tiles_x; t /= g.tiles_x;
    const int tiy = t % g.tiles_y; t /= g.tiles_y;
    const int n = t, y0 = tiy * TC_TH, x0 = tix * TC_TW;

    // ---- stage the haloed A tile, chunk-planar (two 16-byte slots per 8-channel group and pixel)
    for (int it = tid; it < TC_SLOTS * ngroups; it += TC_THREADS) {
      const int grp = it / TC_SLOTS, slot = it - grp * TC_SLOTS;
      const int yy = slot / TC_HW, xx = slot - yy * TC_HW;
      const int gy = y0 + yy - 1, gx = x0 + xx - 1;
      float v[8];
      const int c0 = grp * 8;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W && c0 < Cin) {
        src_load8(p.src, n, gy, gx, c0, min(8, Cin - c0), H, W, v);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      float4* d = reinterpret_cast<float4*>(s_a + ((size_t)(2 * grp) * TC_SLOTS + slot) * 4);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[TC_SLOTS] = make_float4(v[4], v[5], v[6], v[7]);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    __syncthreads();

    // ---- one thread issues the whole K loop: 9 taps x (Cin/8) MMAs of 128 x n_pad x 8
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      uint32_t acc = 0;
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const uint32_t a_tap = a_base + (uint32_t)(ky * TC_HW + kx) * 16;
        for (int kp = 0; kp < ngroups; ++kp) {
          const uint64_t ad = umma_desc(a_tap + (uint32_t)(2 * kp) * TC_PLANE, TC_PLANE, TC_HW * 16);
          const uint64_t bd = umma_desc(w_base + (uint32_t)(tap * nch + 2 * kp) * w_plane, w_plane, 128);
          umma_tf32(tmem_base, ad, bd, g.idesc, acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
    }
    if (!mbar_wait(bar, phase)) {
      if (tid == 0 && status) atomicExch(status, 1);
      break;
    }
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    // ---- epilogue: one output pixel per thread, 16 channels at a time out of TMEM
    const int y = y0 + yl, x = x0 + xl;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < Cout; cb += 16) {
      float a[16];
      tmem_ld16(t_lane + (uint32_t)cb, a);
      const int cn = min(16, Cout - cb);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] += (j < cn) ? __ldg(p.bias + cb + j) : 0.f;
      }
      if (p.epi == CGS_EPI_RELU_POOL) {
        const size_t o = (((size_t)n * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * Cout + cb;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v0 = fmaxf(a[j], 0.f);
          const float v1 = __shfl_xor_sync(0xffffffffu, v0, 1);
          const float v2 = __shfl_xor_sync(0xffffffffu, v0, 8);
          const float v3 = __shfl_xor_sync(0xffffffffu, v0, 9);
          // first maximum in row-major window order wins (ATen max_pool2d)
          float mx = v0; int am = 0;
          if (v1 > mx) { mx = v1; am = 1; }
          if (v2 > mx) { mx = v2; am = 2; }
          if (v3 > mx) { mx = v3; am = 3; }
          if (pool_owner && j < cn) {
            p.out[o + j] = mx;
            if (p.idx_out) p.idx_out[o + j] = (uint8_t)am;
          }
        }
      } else if (p.epi == CGS_EPI_SPLIT_UP) {
        const int C0 = p.C0, C1 = Cout - C0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = cb + j;
          float s = a[j] + __shfl_xor_sync(0xffffffffu, a[j], 1);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          if (j >= cn) continue;
          if (c < C0) {
            if (p.out) p.out[(((size_t)n * H + y) * W + x) * C0 + c] = a[j];
          } else if (p.out2 && pool_owner) {
            p.out2[(((size_t)n * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * C1 + (c - C0)] = s;
          }
        }
      } else {
        const size_t o = (((size_t)n * H + y) * W + x) * Cout + cb;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = a[j];
          if (p.epi == CGS_EPI_LEAKY) v = v > 0.f ? v : v * kLeakySlope;
          else if (p.epi == CGS_EPI_SIGMOID) v = 1.f / (1.f + expf(-v));
          else if (p.epi == CGS_EPI_MUL) v *= (j < cn) ? __ldg(p.mul + o + j) : 0.f;
          a[j] = v;
        }
        if (cn == 16 && (Cout & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(p.out + o + j) = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < cn) p.out[o + j] = a[j];
        }
        if (p.epi == CGS_EPI_SIGMOID && p.idx_out) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < cn) p.idx_out[o + j] = a[j] >= p.thresh ? 1 : 0;
        }
      }
    }
    // TMEM reads and smem operand reads of this tile are complete before the next tile overwrites them
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
  }
}

static int* tc_status_word() {
  static int* d = nullptr;
  if (!d) {
    if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(d, 0, sizeof(int));
  }
  return d;
}

// Whether the tensor-core kernel covers this call; otherwise the caller uses the fp32 FFMA kernel.
bool conv_tc_supported(const cgs_conv3x3_args& a) {
  if (a.H < TC_TH || a.W < TC_TW || (a.H % TC_TH) || (a.W % TC_TW)) return false;
  if (a.epi == CGS_EPI_SPLIT_UP && a.shift2 != 1) return false;
  const int cin_pad = (a.src.C + 7) & ~7, n_pad = (a.Cout + 15) & ~15;
  if (n_pad > 256) return false;
  const size_t smem = (size_t)9 * cin_pad * n_pad * 4 + (size_t)cin_pad * TC_SLOTS * 4 + 64;
  return smem <= 200 * 1024;
}

int launch_conv_tc(const cgs_conv3x3_args& a, cudaStream_t st) {
  TcGeom g;
  g.tiles_x = a.W / TC_TW;
  g.tiles_y = a.H / TC_TH;
  g.ntiles = a.B * g.tiles_x * g.tiles_y;
  g.cin_pad = (a.src.C + 7) & ~7;
  g.nch = g.cin_pad / 4;
  g.n_pad = (a.Cout + 15) & ~15;
  g.tmem_cols = 32;
  while (g.tmem_cols < g.n_pad) g.tmem_cols *= 2;
  // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2), K-major both, N>>3 at 17, M>>4 at 24
  g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.n_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t smem = (size_t)9 * g.cin_pad * g.n_pad * 4 + (size_t)g.cin_pad * TC_SLOTS * 4 + 64;
  static bool attr_done = false;
  static int sms = 148;
  if (!attr_done) {
    cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    attr_done = true;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : per_sm;
  const int tmem_limit = 512 / g.tmem_cols;
  if (per_sm > tmem_limit) per_sm = tmem_limit;
  if (per_sm > 8) per_sm = 8;
  int grid = sms * per_sm;
  if (grid > g.ntiles) grid = g.ntiles;
  conv3x3_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a, g, tc_status_word());
  return check_launch("conv3x3_tc");
}

int conv_tc_status() {
  int* d = tc_status_word();
  int h = 0;
  if (d) cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost);
  return h;
}

}  // namespace cgs

extern "C" int cgs_tc_status(void) { return cgs::conv_tc_status(); }
