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
// split.  Weights are repacked once per CTA into the matching B layout and stay resident.
//
// Pipeline (persistent CTA, double-buffered A tile and TMEM accumulator):
//     stage(i+1) | MMA(i) in flight  ->  wait MMA(i)  ->  issue MMA(i+1)  ->  epilogue(i) | MMA(i+1) in flight
// so global-load latency, tensor-core latency and the epilogue overlap with one __syncthreads per tile.
// Operand prologues (concat+upsample, dropout, pool/ReLU/sigmoid/leaky backward, uint8 cast + roll) are
// the same src_load8 loaders as the fp32 kernel.
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

template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NC]);

template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
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

template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Tight try_wait loop, bounded: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t.reg .u32 cnt;\n\tmov.u32 cnt, 0;\n"
      "TC_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t@P1 bra TC_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\tsetp.lt.u32 P1, cnt, 4000000;\n\t@P1 bra TC_WAIT;\n\t"
      "mov.u32 %0, 0;\n\tbra TC_END;\n"
      "TC_DONE:\n\tmov.u32 %0, 1;\n"
      "TC_END:\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Epilogue for NC channels [cb, cb+NC) of this thread's output pixel.
template <int NC>
__device__ __forceinline__ void tc_epilogue(const cgs_conv3x3_args& p, uint32_t taddr, int cb, int n, int y, int x,
                                            bool odd_x, bool odd_y) {
  const int H = p.H, W = p.W, Cout = p.Cout;
  float a[NC];
  tmem_ld<NC>(taddr + (uint32_t)cb, a);
  const int cn = min(NC, Cout - cb);
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < NC; ++j) a[j] += (j < cn) ? __ldg(p.bias + cb + j) : 0.f;
  }
  if (p.epi == CGS_EPI_RELU_POOL) {
    // 2x2 window = lanes (l, l^1, l^8, l^9); first maximum in row-major order wins (ATen max_pool2d).
    float mx[NC];
    uint32_t abits = 0;   // per channel: 1 if the right element of this row's pair is the (strict) max
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float v = fmaxf(a[j], 0.f);
      const float o = __shfl_xor_sync(0xffffffffu, v, 1);
      const float left = odd_x ? o : v, right = odd_x ? v : o;
      const bool r = right > left;
      mx[j] = r ? right : left;
      abits |= (r ? 1u : 0u) << j;
    }
    const uint32_t obits = __shfl_xor_sync(0xffffffffu, abits, 8);
    const uint32_t tbits = odd_y ? obits : abits, bbits = odd_y ? abits : obits;
    const bool owner = !odd_x && !odd_y;
    float res[NC];
    uint32_t idxw[(NC + 3) / 4];
#pragma unroll
    for (int q = 0; q < (NC + 3) / 4; ++q) idxw[q] = 0;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float o = __shfl_xor_sync(0xffffffffu, mx[j], 8);
      const float top = odd_y ? o : mx[j], bot = odd_y ? mx[j] : o;
      const bool b = bot > top;
      res[j] = b ? bot : top;
      const uint32_t am = b ? (2u + ((bbits >> j) & 1u)) : ((tbits >> j) & 1u);
      idxw[j >> 2] |= am << (8 * (j & 3));
    }
    if (owner) {
      const size_t o = (size_t)(unsigned)((n * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * (unsigned)Cout + cb;
      if (cn == NC && (Cout & 3) == 0) {
#pragma unroll
        for (int j = 0; j < NC; j += 4) {
          *reinterpret_cast<float4*>(p.out + o + j) = make_float4(res[j], res[j + 1], res[j + 2], res[j + 3]);
          if (p.idx_out) *reinterpret_cast<uint32_t*>(p.idx_out + o + j) = idxw[j >> 2];
        }
      } else {
#pragma unroll
        for (int j = 0; j < NC; ++j)
          if (j < cn) {
            p.out[o + j] = res[j];
            if (p.idx_out) p.idx_out[o + j] = (uint8_t)((idxw[j >> 2] >> (8 * (j & 3))) & 0xff);
          }
      }
    }
    return;
  }
  if (p.epi == CGS_EPI_SPLIT_UP) {
    const int C0 = p.C0, C1 = Cout - C0;
    const bool owner = !odd_x && !odd_y;
    const size_t po = (size_t)(unsigned)((n * H + y) * W + x) * (unsigned)C0;
    const size_t pu = (size_t)(unsigned)((n * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * (unsigned)C1;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = cb + j;
      float s = a[j] + __shfl_xor_sync(0xffffffffu, a[j], 1);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      if (j >= cn) continue;
      if (c < C0) {
        if (p.out) p.out[po + c] = a[j];
      } else if (p.out2 && owner) {
        p.out2[pu + (c - C0)] = s;
      }
    }
    return;
  }
  const size_t o = (size_t)(unsigned)((n * H + y) * W + x) * (unsigned)Cout + cb;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float v = a[j];
    if (p.epi == CGS_EPI_LEAKY) v = v > 0.f ? v : v * kLeakySlope;
    else if (p.epi == CGS_EPI_SIGMOID) v = 1.f / (1.f + expf(-v));
    else if (p.epi == CGS_EPI_MUL) v *= (j < cn) ? __ldg(p.mul + o + j) : 0.f;
    a[j] = v;
  }
  if (cn == NC && (Cout & 3) == 0) {
#pragma unroll
    for (int j = 0; j < NC; j += 4) *reinterpret_cast<float4*>(p.out + o + j) = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if (j < cn) p.out[o + j] = a[j];
  }
  if (p.epi == CGS_EPI_SIGMOID && p.idx_out) {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if (j < cn) p.idx_out[o + j] = a[j] >= p.thresh ? 1 : 0;
  }
}

// Optional phase trace (tools/tc_trace.py): CTA 0 / thread 0 records clock64() at the phase boundaries of
// its first 16 tiles.  NULL in production.
__device__ long long* g_tc_trace = nullptr;
#define TC_MARK(k) do { if (trace && tile_no < 16) trace[tile_no * 8 + (k)] = clock64(); } while (0)

__global__ void __launch_bounds__(TC_THREADS) conv3x3_tc_kernel(const cgs_conv3x3_args p, const TcGeom g, int* __restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int H = p.H, W = p.W, Cin = p.src.C, Cout = p.Cout;
  const int nch = g.nch, n_pad = g.n_pad;
  const int a_floats = nch * TC_SLOTS * 4;                                           // one A buffer
  float* s_w = reinterpret_cast<float*>(smem_raw);                                   // [9][nch][n_pad][4]
  float* s_a = s_w + (size_t)9 * nch * n_pad * 4;                                    // [2][nch][180][4]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_a + (size_t)2 * a_floats);
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
    const int rows = 9 * nch * n_pad;
    for (int r = tid; r < rows; r += TC_THREADS) {
      const int co = r % n_pad, tq = r / n_pad;
      const int q = tq % nch, t = tq / nch;
      float v[4];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int ci = q * 4 + c4;
        v[c4] = 0.f;
        if (ci < Cin && co < Cout)
          v[c4] = p.transposed ? __ldg(p.w + ((size_t)ci * Cout + co) * 9 + (8 - t)) : __ldg(p.w + ((size_t)co * Cin + ci) * 9 + t);
      }
      *reinterpret_cast<float4*>(s_w + (size_t)r * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w), bar = smem_u32(s_bar);
  const uint32_t w_plane = (uint32_t)n_pad * 16;     // bytes between consecutive 4-channel weight planes
  const int ngroups = g.cin_pad >> 3;
  const int tpf = g.tiles_x * g.tiles_y;

  // this thread's output pixel inside a tile: TMEM lane m = tid -> (yl, xl) = (m / 8, m % 8)
  const int yl = tid >> 3, xl = tid & 7;
  const bool odd_x = xl & 1, odd_y = yl & 1;
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);

  auto tile_origin = [&](int tile, int& n, int& y0, int& x0) {
    n = tile / tpf;
    const int r = tile - n * tpf;
    const int tiy = r / g.tiles_x;
    y0 = tiy * TC_TH;
    x0 = (r - tiy * g.tiles_x) * TC_TW;
  };
  auto stage = [&](int tile, int buf) {
    int n, y0, x0;
    tile_origin(tile, n, y0, x0);
    float* sa = s_a + (size_t)buf * a_floats;
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
      float4* d = reinterpret_cast<float4*>(sa + ((size_t)(2 * grp) * TC_SLOTS + slot) * 4);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[TC_SLOTS] = make_float4(v[4], v[5], v[6], v[7]);
    }
  };
  auto issue = [&](int buf) {   // one thread: 9 taps x (Cin/8) MMAs of 128 x n_pad x 8, then commit
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t ab = a_base + (uint32_t)buf * (uint32_t)a_floats * 4u;
    const uint32_t td = tmem_base + (uint32_t)buf * (uint32_t)n_pad;
    uint32_t acc = 0;
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap - ky * 3;
      const uint32_t a_tap = ab + (uint32_t)(ky * TC_HW + kx) * 16;
      for (int kp = 0; kp < ngroups; ++kp) {
        const uint64_t ad = umma_desc(a_tap + (uint32_t)(2 * kp) * TC_PLANE, TC_PLANE, TC_HW * 16);
        const uint64_t bd = umma_desc(w_base + (uint32_t)(tap * nch + 2 * kp) * w_plane, w_plane, 128);
        umma_tf32(td, ad, bd, g.idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
  };

  long long* trace = (blockIdx.x == 0 && threadIdx.x == 0) ? g_tc_trace : nullptr;
  int tile_no = 0;
  int tile = blockIdx.x;
  if (tile < g.ntiles) {
    stage(tile, 0);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) issue(0);
  }
  uint32_t phase = 0;
  int buf = 0;
  for (; tile < g.ntiles; tile += gridDim.x, buf ^= 1, ++tile_no) {
    const int next = tile + gridDim.x;
    const bool has_next = next < g.ntiles;
    TC_MARK(0);
    if (has_next) stage(next, buf ^ 1);                    // overlaps with MMA(tile)
    TC_MARK(1);
    if (!mbar_wait(bar, phase)) {
      if (tid == 0 && status) atomicExch(status, 1);
      break;
    }
    phase ^= 1;
    TC_MARK(2);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();   // next tile staged by all; everyone is done reading TMEM[buf^1] (epilogue of tile-1)
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    TC_MARK(3);
    if (has_next && tid == 0) issue(buf ^ 1);              // MMA(next) overlaps with the epilogue below
    TC_MARK(4);

    int n, y0, x0;
    tile_origin(tile, n, y0, x0);
    const int y = y0 + yl, x = x0 + xl;
    const uint32_t taddr = t_lane + (uint32_t)buf * (uint32_t)n_pad;
    if (Cout <= 8) {
      tc_epilogue<8>(p, taddr, 0, n, y, x, odd_x, odd_y);
    } else {
      for (int cb = 0; cb < Cout; cb += 16) tc_epilogue<16>(p, taddr, cb, n, y, x, odd_x, odd_y);
    }
    TC_MARK(5);
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
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

static size_t tc_smem_bytes(int cin_pad, int n_pad) {
  return (size_t)9 * cin_pad * n_pad * 4 + (size_t)2 * cin_pad * TC_SLOTS * 4 + 64;
}

// Whether the tensor-core kernel covers this call; otherwise the caller uses the fp32 FFMA kernel.
bool conv_tc_supported(const cgs_conv3x3_args& a) {
  if (a.H < TC_TH || a.W < TC_TW || (a.H % TC_TH) || (a.W % TC_TW)) return false;
  if (a.epi == CGS_EPI_SPLIT_UP && a.shift2 != 1) return false;
  // Cin <= 4 (the RGB input layer): K = 27 would be padded to 72 and the implicit GEMM re-reads the tile 9x through
  // the tensor core's operand path; the register-blocked FFMA kernel is faster there (profiles/README.md).
  if (a.src.C <= 4) return false;
  const int cin_pad = (a.src.C + 7) & ~7, n_pad = (a.Cout + 15) & ~15;
  if (2 * n_pad > 512) return false;
  return tc_smem_bytes(cin_pad, n_pad) <= 200 * 1024;
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
  while (g.tmem_cols < 2 * g.n_pad) g.tmem_cols *= 2;   // double-buffered accumulator
  // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2), K-major both, N>>3 at 17, M>>4 at 24
  g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.n_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t smem = tc_smem_bytes(g.cin_pad, g.n_pad);
  cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int sms = device_sms();
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : per_sm;
  const int tmem_limit = 512 / g.tmem_cols;
  if (per_sm > tmem_limit) per_sm = tmem_limit;
  if (per_sm > 6) per_sm = 6;
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

// Debug: point the phase trace at a device buffer of 16*8 int64 (NULL disables).  Not part of the product API.
extern "C" int cgs_tc_set_trace(long long* dev_buf) {
  return cudaMemcpyToSymbol(cgs::g_tc_trace, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -2;
}
