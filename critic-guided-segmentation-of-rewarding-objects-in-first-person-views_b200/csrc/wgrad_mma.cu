// Weight gradient of the 3x3 convolutions on the tensor cores (TF32 operands, fp32 accumulate).
//
// GEMM view:  dW[(ci,tap), co] = sum_pixels  X[pixel + tap, ci] * dY[pixel, co]   (+ a row of ones -> bias grad)
// i.e. M = 9*Cin (tiny), N = Cout (tiny), K = B*H*W (huge): the opposite of what tcgen05 wants.  A tcgen05
// A-operand must sit in shared memory in the canonical core-matrix layout, so the nine shifted views of the
// haloed tile would have to be materialised as a 9x im2col copy; the warp-level mma.m16n8k8 instead takes its
// A fragment from registers, so every lane GATHERS its two (ci,tap) rows straight from the planar haloed
// tile (row base = ci*plane + ky*row + kx, then +x along the reduction).  That makes the instruction mix
// 22 LDS + 5 HMMA per 8 pixels per warp instead of ~170 LDS + 1150 FFMA, and the kernel staging-bound.
// The tcgen05 kernels (conv_tc.cu) keep fprop and dgrad, where M = pixels is the large dimension.
//
// CTA = one 8-channel (ci) block x one 8-channel (co) block x a pixel tile; the 8 warps split the tile rows
// (split-K), partial tiles are reduced through shared memory and pushed with one RED per weight per CTA.
#include "common.cuh"

namespace cgs {


constexpr int WM_CI = 8, WM_CO = 8;   // channel block; M tiles: 5 x m16 = 80 rows >= 72 (ci,tap) rows + 1 bias row,
                                      // or 2 x m16 = 32 rows >= 27 + 1 when Cin <= 3 (the RGB input layer)

__device__ __forceinline__ uint32_t f2tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(f));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float tf32r(float f) { return __uint_as_float(f2tf32(f)); }

template <int WM_MT>
__global__ void __launch_bounds__(256) wgrad3x3_mma_kernel(const cgs_wgrad3x3_args p, const WgGeom g) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int H = p.H, W = p.W, Cin = p.x.C, Cout = p.dy.C;
  const int th = g.th, tw = g.tw, sh = th + 2, sw = tw + 2;
  float* s_x = smem;                                   // [WM_CI][fpc*sh rows][rsx]
  float* s_y = smem + (size_t)g.xplanes * g.psx;       // [WM_CO][fpc*th rows][rsy]

  const int co0 = blockIdx.y * WM_CO, ci0 = blockIdx.z * WM_CI;
  const int con = min(WM_CO, Cout - co0), cin = min(WM_CI, Cin - ci0);

  // ---- per-lane A rows: m = ci*9 + tap (the OIHW order of dw), m == 72 is the all-ones bias row
  const int warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  // branch-free gather: value = x[off] * mulA + addA with (1,0) data row, (0,1) ones row, (0,0) pad row
  // WM_MT == 5 (8 input channels per block): rows of M-tile mt are (tap 2mt | tap 2mt+1) x ci, row 8 of mt 4 = ones (bias);
  // the mma's k index maps to pixels as k = t -> 2t, k = t+4 -> 2t+1, so the 4 consecutive pixels 2t..2t+3 of a filter row
  // (two 64-bit loads from the planar tile) feed all three kx taps: 6 LDS.64 per 8 pixels instead of 20 LDS.32 + 20 FMA.
  // WM_MT == 2 (RGB layer): rows m = ci*9 + tap gathered one by one, as before.
  int offA[WM_MT][2];
  float mulA[WM_MT][2], addA[WM_MT][2];
#pragma unroll
  for (int mt = 0; mt < WM_MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int mrow = 16 * mt + gid + 8 * h;
      int off = 0;
      float mu = 0.f, ad = 0.f;
      if (mrow < 9 * cin) {
        const int ci = mrow / 9, tap = mrow - ci * 9;
        off = ci * g.psx + (tap / 3) * g.rsx + (tap % 3);
        mu = 1.f;
      } else if (mrow == 9 * cin) {
        ad = 1.f;
      }
      offA[mt][h] = off; mulA[mt][h] = mu; addA[mt][h] = ad;
    }
  float acc[WM_MT][4];
#pragma unroll
  for (int mt = 0; mt < WM_MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[mt][j] = 0.f;

  // ---- persistent over pixel tiles: partial sums stay in registers, ONE reduction + RED round per CTA at the end
  //      (same-address REDs serialise in L2 at ~50 ns each: the number of CTAs, not the work, set the run time)
  for (int bid0 = blockIdx.x; bid0 < g.nblk; bid0 += gridDim.x) {
  int bid = bid0;
  const int tix = bid % g.tiles_x; bid /= g.tiles_x;
  const int tiy = bid % g.tiles_y; bid /= g.tiles_y;
  const int n0 = bid * g.fpc, y0 = tiy * th, x0 = tix * tw;
  __syncthreads();      // the previous tile's MMA loop is done with the staged tiles

  // ---- stage X (haloed) and dY tiles, planar [channel][row][x]
  {
    const int npix = g.fpc * sh * sw;
#pragma unroll 2
    for (int pix = tid; pix < npix; pix += 256) {
      const int row = fdiv(pix, g.dsw), xx = pix - row * sw;
      const int ff = fdiv(row, g.dsh), yy = row - ff * sh;
      const int gy = y0 + yy - 1, gx = x0 + xx - 1, nn = n0 + ff;
      float v[8];
      if (nn < p.B && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        src_load8(p.x, nn, gy, gx, ci0, cin, H, W, v);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      float* d = s_x + row * g.rsx + xx;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < cin) d[i * g.psx] = tf32r(v[i]);      // round once here, not 9x in the MMA loop
      if (cin > 4) {
#pragma unroll
        for (int i = 4; i < 8; ++i)
          if (i < cin) d[i * g.psx] = tf32r(v[i]);
      }
    }
    if (p.dy.mode == CGS_SRC_POOLBWD && (p.dy.C & 7) == 0) {
      // ReLU + max-pool backward at POOLED granularity: one (dE, E, argmax) load per pooled element, scattered to the
      // arg-max position of its 2x2 window (the other three get zero) -- 4x fewer loads than per output pixel.
      const int hp = th >> 1, wp = tw >> 1, H2 = H >> 1, W2 = W >> 1;
      const int npool = g.fpc * hp * wp;
#pragma unroll 2
      for (int pp = tid; pp < npool; pp += 256) {
        const int r2 = fdiv(pp, g.dwp), xx2 = pp - r2 * wp;
        const int ff = fdiv(r2, g.dhp), yy2 = r2 - ff * hp;
        const int nn = n0 + ff;
        float de[8], e[8];
        unsigned long long ib = 0;
        if (nn < p.B) {
          const size_t o = (size_t)(unsigned)((nn * H2 + (y0 >> 1) + yy2) * W2 + (x0 >> 1) + xx2) * (unsigned)p.dy.C + co0;
          load8(p.dy.a + o, 8, true, de);
          load8(p.dy.b + o, 8, true, e);
          ib = __ldg(reinterpret_cast<const unsigned long long*>(p.dy.idx + o));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { de[i] = 0.f; e[i] = 0.f; }
        }
        float* d = s_y + (ff * th + 2 * yy2) * g.rsy + 2 * xx2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pos = (int)((ib >> (8 * i)) & 0xff);
          const float v = (e[i] > 0.f) ? tf32r(de[i]) : 0.f;
          float* q = d + i * g.psy;
          q[0] = pos == 0 ? v : 0.f;
          q[1] = pos == 1 ? v : 0.f;
          q[g.rsy] = pos == 2 ? v : 0.f;
          q[g.rsy + 1] = pos == 3 ? v : 0.f;
        }
      }
    } else {
    const int npy = g.fpc * th * tw;
#pragma unroll 2
    for (int pix = tid; pix < npy; pix += 256) {
      const int row = fdiv(pix, g.dtw), xx = pix - row * tw;
      const int ff = fdiv(row, g.dth), yy = row - ff * th;
      const int gy = y0 + yy, gx = x0 + xx, nn = n0 + ff;
      float v[8];
      if (nn < p.B) {
        src_load8(p.dy, nn, gy, gx, co0, con, H, W, v);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      float* d = s_y + row * g.rsy + xx;
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i * g.psy] = (i < con) ? tf32r(v[i]) : 0.f;
    }
    }
  }
  __syncthreads();

  const int R = g.fpc * th;
  if (WM_MT == 5) {
    const uint32_t ones = gid == 0 ? 0x3f800000u : 0u;
    for (int r = warp; r < R; r += 8) {
      const int ff = fdiv(r, g.dth), yy = r - ff * th;
      const float* xrow = s_x + gid * g.psx + (ff * sh + yy) * g.rsx + 2 * tig;
      const float* yrow = s_y + gid * g.psy + r * g.rsy + 2 * tig;
      for (int xb = 0; xb < tw; xb += 8) {
        const float2 bb = *reinterpret_cast<const float2*>(yrow + xb);
        const uint32_t b0 = __float_as_uint(bb.x), b1 = __float_as_uint(bb.y);
        uint32_t v[3][4];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float2 lo = *reinterpret_cast<const float2*>(xrow + ky * g.rsx + xb);
          const float2 hi = *reinterpret_cast<const float2*>(xrow + ky * g.rsx + xb + 2);
          v[ky][0] = __float_as_uint(lo.x); v[ky][1] = __float_as_uint(lo.y);
          v[ky][2] = __float_as_uint(hi.x); v[ky][3] = __float_as_uint(hi.y);
        }
#pragma unroll
        for (int mt = 0; mt < WM_MT; ++mt) {
          const int ta = (2 * mt) % 9, tb = (2 * mt + 1) % 9;       // (% 9 only keeps the WM_MT == 2 instantiation in bounds)
          const uint32_t a0 = v[ta / 3][ta % 3], a2 = v[ta / 3][ta % 3 + 1];
          const uint32_t a1 = mt < 4 ? v[tb / 3][tb % 3] : ones, a3 = mt < 4 ? v[tb / 3][tb % 3 + 1] : ones;
          mma_tf32(acc[mt], a0, a1, a2, a3, b0, b1);
        }
      }
    }
  } else {
  for (int r = warp; r < R; r += 8) {
    const int ff = fdiv(r, g.dth), yy = r - ff * th;
    const float* xrow = s_x + (ff * sh + yy) * g.rsx + tig;
    const float* yrow = s_y + gid * g.psy + r * g.rsy + tig;
    for (int xb = 0; xb < tw; xb += 8) {
      const uint32_t b0 = __float_as_uint(yrow[xb]), b1 = __float_as_uint(yrow[xb + 4]);
#pragma unroll
      for (int mt = 0; mt < WM_MT; ++mt) {
        uint32_t a[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int off = offA[mt][h];
          a[h] = __float_as_uint(fmaf(xrow[off + xb], mulA[mt][h], addA[mt][h]));
          a[2 + h] = __float_as_uint(fmaf(xrow[off + xb + 4], mulA[mt][h], addA[mt][h]));
        }
        mma_tf32(acc[mt], a[0], a[1], a[2], a[3], b0, b1);
      }
    }
  }
  }

  }   // tile loop

  // ---- reduce the 8 warps' partial tiles through shared memory (reuse the X tile), then one RED per weight
  __syncthreads();
  float* s_red = smem;   // [8 warps][80 rows][8 cols]
#pragma unroll
  for (int mt = 0; mt < WM_MT; ++mt) {
    float* d = s_red + (warp * 16 * WM_MT + 16 * mt + gid) * 8 + 2 * tig;
    d[0] = acc[mt][0]; d[1] = acc[mt][1];
    d[64] = acc[mt][2]; d[65] = acc[mt][3];     // row + 8
  }
  __syncthreads();
  for (int e = tid; e < 16 * WM_MT * 8; e += 256) {
    const int mrow = e >> 3, n = e & 7;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w * 16 * WM_MT * 8 + e];
    if (n >= con) continue;
    if (WM_MT == 5) {
      const int tap = 2 * (mrow >> 4) + ((mrow >> 3) & 1), ci = mrow & 7;
      if (tap < 9) {
        if (ci < cin) atomicAdd(p.dw + ((size_t)(co0 + n) * Cin + ci0 + ci) * 9 + tap, s);
      } else if (tap == 9 && ci == 0 && p.db && blockIdx.z == 0) {
        atomicAdd(p.db + co0 + n, s);
      }
    } else if (mrow < 9 * cin) {
      atomicAdd(p.dw + ((size_t)(co0 + n) * Cin + ci0) * 9 + mrow, s);
    } else if (mrow == 9 * cin && p.db && blockIdx.z == 0) {
      atomicAdd(p.db + co0 + n, s);
    }
  }
}

bool wgrad_mma_supported(const cgs_wgrad3x3_args& a) { return a.W >= 8 && a.H >= 8; }

int launch_wgrad_mma(const cgs_wgrad3x3_args& a, cudaStream_t st) {
  WgGeom g;
  g.th = a.H < 32 ? a.H : 32;
  g.tw = a.W < 32 ? a.W : 32;
  if (g.th * g.tw > 512 && (long)a.B * (a.H / g.th) * (a.W / g.tw) < 592) g.th /= 2;   // more CTAs when the batch is small
  g.tiles_y = a.H / g.th; g.tiles_x = a.W / g.tw;
  g.fpc = 1;
  while (g.fpc * g.th * g.tw < 512 && (long)((a.B + 2 * g.fpc - 1) / (2 * g.fpc)) >= 148) g.fpc *= 2;
  // A-fragment gathers read, per instruction, 8 (ci,tap) rows x 4 consecutive pixels: with a row stride == 8 (mod 16)
  // the three filter rows land in disjoint bank ranges; dY planes 4 banks apart make the B-fragment reads conflict-free.
  g.rsx = (g.tw + 2 + 7) & ~7; if ((g.rsx & 15) == 0) g.rsx += 8;
  g.rsy = g.tw;
  g.psx = g.fpc * (g.th + 2) * g.rsx; g.psx += (56 - (g.psx % 32)) % 32;   // == 24 (mod 32)
  g.psy = g.fpc * g.th * g.rsy;
  if (a.x.C <= 3) g.psy += (36 - (g.psy % 32)) % 32;   // == 4 (mod 32): 32-bit B-fragment loads, pixels t and t+4
  else g.psy += (40 - (g.psy % 32)) % 32;              // == 8 (mod 32): 64-bit loads of the pixel pair (2t, 2t+1)
  g.dsw = make_fastdiv(g.tw + 2); g.dsh = make_fastdiv(g.th + 2);
  g.dtw = make_fastdiv(g.tw);     g.dth = make_fastdiv(g.th);
  g.dwp = make_fastdiv(g.tw / 2); g.dhp = make_fastdiv(g.th / 2);
  g.xplanes = a.x.C < WM_CI ? a.x.C : WM_CI;
  size_t smem = ((size_t)g.xplanes * g.psx + (size_t)WM_CO * g.psy) * sizeof(float);
  const size_t red = (size_t)8 * 16 * 5 * 8 * sizeof(float);
  if (smem < red) smem = red;
  cudaFuncSetAttribute(wgrad3x3_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(wgrad3x3_mma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  g.nblk = ((a.B + g.fpc - 1) / g.fpc) * g.tiles_x * g.tiles_y;
  const int gy = (a.dy.C + WM_CO - 1) / WM_CO, gz = (a.x.C + WM_CI - 1) / WM_CI;
  const int sms = device_sms();
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  int gx = (sms * per_sm) / (gy * gz);
  gx = gx < 1 ? 1 : gx;
  if (gx > g.nblk) gx = g.nblk;
  dim3 grid(gx, gy, gz);
  if (a.x.C <= 3) wgrad3x3_mma_kernel<2><<<grid, 256, smem, st>>>(a, g);
  else wgrad3x3_mma_kernel<5><<<grid, 256, smem, st>>>(a, g);
  return check_launch("wgrad3x3_mma");
}

}  // namespace cgs
