// fp32 (CUDA-core FFMA) 3x3 stride-1 pad-1 convolution: fprop / dgrad / wgrad with fused
// prologues (concat + nearest upsample, dropout mask, ReLU+max-pool / sigmoid / leaky
// backward) and epilogues (bias, ReLU + 2x2 max-pool + argmax, leaky, sigmoid + threshold,
// dropout-mask multiply, concat/upsample backward split).  This is the exact-fp32 parity
// path (rtol 1e-4 vs the reference); the tcgen05 TF32 path lives in conv_tc.cu.
//
// Layout: activations NHWC fp32, weights OIHW fp32 (reference state_dict layout).
// Tiling: a CTA owns `fpc` frames x (2*tph x 2*tpw) output pixels x CO_T output channels;
// each thread owns a 2x2 pixel patch (== one max-pool window) x CO_T channels in registers.
// Input channels are streamed through shared memory in chunks of CI_T planes with a 1-px halo.
#include "common.cuh"

namespace cgs {

constexpr int CI_T = 8;

struct ConvGeom {
  int tph, tpw, fpc;      // patches per tile (rows, cols), frames per CTA
  int tiles_y, tiles_x;   // tiles per frame
  int rs, ps;             // smem row stride / plane stride (floats)
  FastDiv dsw, dsh;       // dividers by the haloed tile width / height
};

template <int CO_T>
__device__ __forceinline__ void load_w(const float* wp, float (&w)[CO_T]) {
  if constexpr (CO_T == 1) {
    w[0] = wp[0];
  } else {
#pragma unroll
    for (int i = 0; i < CO_T / 4; ++i) {
      float4 t = reinterpret_cast<const float4*>(wp)[i];
      w[4 * i + 0] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
    }
  }
}

template <int CO_T>
__global__ void __launch_bounds__(256, (CO_T <= 8 ? 3 : 2)) conv3x3_kernel(const cgs_conv3x3_args p, const ConvGeom g) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int H = p.H, W = p.W, Cin = p.src.C, Cout = p.Cout;
  const int th = 2 * g.tph, tw = 2 * g.tpw, sh = th + 2, sw = tw + 2;
  float* s_in = smem;                                // [fpc][CI_T][ps]
  float* s_w = smem + (size_t)g.fpc * CI_T * g.ps;   // [CI_T][9][CO_T]

  int bid = blockIdx.x;
  const int tix = bid % g.tiles_x; bid /= g.tiles_x;
  const int tiy = bid % g.tiles_y; bid /= g.tiles_y;
  const int n0 = bid * g.fpc;
  const int y0 = tiy * th, x0 = tix * tw;
  const int co0 = blockIdx.y * CO_T;

  const int ppf = g.tph * g.tpw;
  const int f = tid / ppf, ty = (tid % ppf) / g.tpw, tx = tid % g.tpw;
  const int n = n0 + f;
  const bool active = (f < g.fpc) && (n < p.B);

  float acc[4][CO_T];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CO_T; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CI_T) {
    const int ci_n = min(CI_T, Cin - c0);
    __syncthreads();
    // ---- stage the haloed input chunk: one pixel (<= 8 channels, 128-bit loads) per thread iteration
    const int npix = g.fpc * sh * sw;
#pragma unroll 1
    for (int pix = tid; pix < npix; pix += nthr) {
      const int row = fdiv(pix, g.dsw), xx = pix - row * sw;
      const int ff = fdiv(row, g.dsh), yy = row - ff * sh;
      const int gy = y0 + yy - 1, gx = x0 + xx - 1, nn = n0 + ff;
      float v[8];
      if (nn < p.B && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        src_load8(p.src, nn, gy, gx, c0, ci_n, H, W, v);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      float* d = s_in + (ff * CI_T) * g.ps + yy * g.rs + xx;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < ci_n) d[i * g.ps] = v[i];
      if (ci_n > 4) {
#pragma unroll
        for (int i = 4; i < 8; ++i)
          if (i < ci_n) d[i * g.ps] = v[i];
      }
    }
    // ---- stage the weight chunk as [ci][tap][co]
#pragma unroll 4
    for (int e = tid; e < ci_n * 9 * CO_T; e += nthr) {
      const int co = e % CO_T, t = (e / CO_T) % 9, ci = e / (9 * CO_T);
      const int gco = co0 + co;
      float v = 0.f;
      if (gco < Cout)
        v = p.transposed ? __ldg(p.w + ((size_t)(c0 + ci) * Cout + gco) * 9 + (8 - t))
                         : __ldg(p.w + ((size_t)gco * Cin + (c0 + ci)) * 9 + t);
      s_w[e] = v;
    }
    __syncthreads();
    if (active) {
      for (int ci = 0; ci < ci_n; ++ci) {
        const float* ip = s_in + (f * CI_T + ci) * g.ps + (2 * ty) * g.rs + 2 * tx;
        float in[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float2 a = *reinterpret_cast<const float2*>(ip + r * g.rs);
          float2 b = *reinterpret_cast<const float2*>(ip + r * g.rs + 2);
          in[r][0] = a.x; in[r][1] = a.y; in[r][2] = b.x; in[r][3] = b.y;
        }
        const float* wp = s_w + ci * 9 * CO_T;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float w[CO_T];
          load_w<CO_T>(wp + t * CO_T, w);
          const int ky = t / 3, kx = t % 3;
#pragma unroll
          for (int j = 0; j < CO_T; ++j) {
            acc[0][j] = fmaf(in[ky][kx], w[j], acc[0][j]);
            acc[1][j] = fmaf(in[ky][kx + 1], w[j], acc[1][j]);
            acc[2][j] = fmaf(in[ky + 1][kx], w[j], acc[2][j]);
            acc[3][j] = fmaf(in[ky + 1][kx + 1], w[j], acc[3][j]);
          }
        }
      }
    }
  }
  if (!active) return;

  // ---------------- epilogue
  const int py = y0 + 2 * ty, px = x0 + 2 * tx;   // top-left pixel of the patch
  if (py >= H || px >= W) return;
  const int con = min(CO_T, Cout - co0);
  float bias[CO_T];
#pragma unroll
  for (int j = 0; j < CO_T; ++j) bias[j] = (p.bias && j < con) ? __ldg(p.bias + co0 + j) : 0.f;

  if (p.epi == CGS_EPI_RELU_POOL) {
    const int h2 = H >> 1, w2 = W >> 1;
    const size_t o = (((size_t)n * h2 + (py >> 1)) * w2 + (px >> 1)) * Cout + co0;
    float mv[CO_T];
    unsigned char am[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
      // ATen max-pool keeps the FIRST maximum in row-major window order (strict >).
      float m = fmaxf(acc[0][j] + bias[j], 0.f);
      int a = 0;
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        float v = fmaxf(acc[q][j] + bias[j], 0.f);
        if (v > m) { m = v; a = q; }
      }
      mv[j] = m; am[j] = (unsigned char)a;
    }
    if constexpr (CO_T % 4 == 0) {
      if (con == CO_T && (Cout & 3) == 0) {
#pragma unroll
        for (int j = 0; j < CO_T; j += 4) {
          *reinterpret_cast<float4*>(p.out + o + j) = make_float4(mv[j], mv[j + 1], mv[j + 2], mv[j + 3]);
          if (p.idx_out)
            *reinterpret_cast<uchar4*>(p.idx_out + o + j) = make_uchar4(am[j], am[j + 1], am[j + 2], am[j + 3]);
        }
        return;
      }
    }
#pragma unroll
    for (int j = 0; j < CO_T; ++j)
      if (j < con) {
        p.out[o + j] = mv[j];
        if (p.idx_out) p.idx_out[o + j] = am[j];
      }
    return;
  }
  if (p.epi == CGS_EPI_SPLIT_UP) {
    // channels [0,C0) -> out (skip gradient); [C0,Cout) -> window sum into out2 (upsample backward)
    const int C0 = p.C0, C1 = Cout - C0, s2 = p.shift2;
    const int h2 = H >> s2, w2 = W >> s2;
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
      const int c = co0 + j;
      if (j >= con) continue;
      if (c < C0) {
        if (p.out) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            p.out[(((size_t)n * H + py + (q >> 1)) * W + px + (q & 1)) * C0 + c] = acc[q][j];
        }
      } else if (p.out2) {
        const float s = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
        float* dst = p.out2 + (((size_t)n * h2 + (py >> s2)) * w2 + (px >> s2)) * C1 + (c - C0);
        if (s2 == 1) *dst = s; else atomicAdd(dst, s);
      }
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const size_t o = (((size_t)n * H + py + (q >> 1)) * W + px + (q & 1)) * Cout + co0;
    float r[CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
      float v = acc[q][j] + bias[j];
      if (p.epi == CGS_EPI_LEAKY) v = v > 0.f ? v : v * kLeakySlope;
      else if (p.epi == CGS_EPI_SIGMOID) v = 1.f / (1.f + expf(-v));
      else if (p.epi == CGS_EPI_MUL) v *= (j < con) ? __ldg(p.mul + o + j) : 0.f;
      r[j] = v;
    }
    bool done = false;
    if constexpr (CO_T % 4 == 0) {
      if (con == CO_T && (Cout & 3) == 0) {
#pragma unroll
        for (int j = 0; j < CO_T; j += 4)
          *reinterpret_cast<float4*>(p.out + o + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        done = true;
      }
    }
    if (!done) {
#pragma unroll
      for (int j = 0; j < CO_T; ++j)
        if (j < con) p.out[o + j] = r[j];
    }
    if (p.epi == CGS_EPI_SIGMOID && p.idx_out) {
#pragma unroll
      for (int j = 0; j < CO_T; ++j)
        if (j < con) p.idx_out[o + j] = r[j] >= p.thresh ? 1 : 0;
    }
  }
}

static void pick_geom(int B, int H, int W, ConvGeom& g, int& nthr, int& nblk) {
  // patches per frame dim
  const int phh = H / 2, pww = W / 2;
  g.tph = phh < 16 ? phh : 16;
  g.tpw = pww < 16 ? pww : 16;
  g.tiles_y = phh / g.tph;
  g.tiles_x = pww / g.tpw;
  int ppf = g.tph * g.tpw;          // threads per frame-tile
  // frames per CTA: fill 128 threads for small maps but keep enough CTAs to cover the SMs
  g.fpc = 1;
  while (ppf * g.fpc < 128 && 2 * g.fpc <= B) g.fpc *= 2;   // >= 4 warps per CTA: these layers are latency-bound
  nthr = ((ppf * g.fpc + 31) / 32) * 32;
  const int sw = 2 * g.tpw + 2, sh = 2 * g.tph + 2;
  g.rs = (sw + 1) & ~1;
  g.ps = sh * g.rs;
  g.ps += (36 - (g.ps % 32)) % 32;  // plane stride == 4 (mod 32)
  g.dsw = make_fastdiv(sw);
  g.dsh = make_fastdiv(sh);
  nblk = ((B + g.fpc - 1) / g.fpc) * g.tiles_x * g.tiles_y;
}

template <int CO_T>
static int launch_conv(const cgs_conv3x3_args& a, cudaStream_t st) {
  ConvGeom g; int nthr, nblk;
  pick_geom(a.B, a.H, a.W, g, nthr, nblk);
  size_t smem = ((size_t)g.fpc * CI_T * g.ps + CI_T * 9 * CO_T) * sizeof(float);
  cudaFuncSetAttribute(conv3x3_kernel<CO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  dim3 grid(nblk, (a.Cout + CO_T - 1) / CO_T);
  conv3x3_kernel<CO_T><<<grid, nthr, smem, st>>>(a, g);
  return check_launch("conv3x3");
}

bool wgrad_pipe_supported(const cgs_wgrad3x3_args& a);
int launch_wgrad_pipe(const cgs_wgrad3x3_args& a, cudaStream_t st);
bool wgrad_mma_supported(const cgs_wgrad3x3_args& a);
int launch_wgrad_mma(const cgs_wgrad3x3_args& a, cudaStream_t st);
bool conv_tc_supported(const cgs_conv3x3_args& a);
int launch_conv_tc(const cgs_conv3x3_args& a, cudaStream_t st);

static int check_src(const cgs_src& s, const char* who) {
  CGS_REQUIRE(s.a != nullptr && s.C > 0, "%s: null operand / C<=0", who);
  if (s.mode == CGS_SRC_CATUP) CGS_REQUIRE(s.b && s.C0 > 0 && s.C0 < s.C && (s.shift == 1 || s.shift == 2), "%s: bad CATUP operand", who);
  if (s.mode == CGS_SRC_POOLBWD) CGS_REQUIRE(s.b && s.idx, "%s: POOLBWD needs E and idx", who);
  if (s.mode == CGS_SRC_SIGGRAD || s.mode == CGS_SRC_LEAKYGRAD) CGS_REQUIRE(s.b, "%s: grad mode needs forward output", who);
  CGS_REQUIRE(s.mode >= 0 && s.mode <= CGS_SRC_U8ROLL, "%s: unknown src mode %d", who, s.mode);
  return 0;
}

// ======================================================================= wgrad
// dW[co][ci][t] += sum_{n,y,x} dY[n,y,x,co] * X[n,y+ky-1,x+kx-1,ci];  dB[co] += sum dY.
// CTA: one (co block, ci block) x a pixel tile; 16 "pixel groups" (rows) x 16 (co,ci)
// register blocks of CO_R x CI_R x 9 taps; rows reduced with warp shuffles, one atomic per
// weight per CTA.

template <int CO_R, int CI_R, int NCO, int NCI>
__global__ void __launch_bounds__(256) wgrad3x3_kernel(const cgs_wgrad3x3_args p, const WgGeom g) {
  static_assert(NCO * NCI == 16, "16 register blocks per CTA");
  constexpr int CO_B = CO_R * NCO, CI_B = CI_R * NCI;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int H = p.H, W = p.W, Cin = p.x.C, Cout = p.dy.C;
  const int th = g.th, tw = g.tw, sh = th + 2, sw = tw + 2;
  float* s_x = smem;                                   // [CI_B][fpc*sh rows][rsx]
  float* s_y = smem + (size_t)CI_B * g.psx;            // [CO_B][fpc*th rows][rsy]

  int bid = blockIdx.x;
  const int tix = bid % g.tiles_x; bid /= g.tiles_x;
  const int tiy = bid % g.tiles_y; bid /= g.tiles_y;
  const int n0 = bid * g.fpc, y0 = tiy * th, x0 = tix * tw;
  const int co0 = blockIdx.y * CO_B, ci0 = blockIdx.z * CI_B;
  const int con = min(CO_B, Cout - co0), cin = min(CI_B, Cin - ci0);

  // ---- stage X (haloed) and dY tiles: one pixel x <=8 channels per thread iteration
  {
    const int npix = g.fpc * sh * sw;
    for (int cb = 0; cb < cin; cb += 8) {
      const int cn = min(8, cin - cb);
#pragma unroll 2
      for (int pix = tid; pix < npix; pix += 256) {
        const int row = fdiv(pix, g.dsw), xx = pix - row * sw;
        const int ff = fdiv(row, g.dsh), yy = row - ff * sh;
        const int gy = y0 + yy - 1, gx = x0 + xx - 1, nn = n0 + ff;
        float v[8];
        if (nn < p.B && gy >= 0 && gy < H && gx >= 0 && gx < W) {
          src_load8(p.x, nn, gy, gx, ci0 + cb, cn, H, W, v);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        float* d = s_x + cb * g.psx + row * g.rsx + xx;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < cn) d[i * g.psx] = v[i];
      }
    }
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
      for (int i = 0; i < 8; ++i)
        if (i < con) d[i * g.psy] = v[i];
    }
  }
  __syncthreads();

  const int pg = tid & 15, cc = tid >> 4;
  const int cco = cc % NCO, cci = cc / NCO;
  const int col = cco * CO_R, cil = cci * CI_R;   // local channel offsets
  const bool live = (col < con) && (cil < cin);

  float acc[CO_R][CI_R][9];
  float bsum[CO_R];
#pragma unroll
  for (int a = 0; a < CO_R; ++a) {
    bsum[a] = 0.f;
#pragma unroll
    for (int b = 0; b < CI_R; ++b)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[a][b][t] = 0.f;
  }
  const bool do_bias = p.db != nullptr && blockIdx.z == 0 && cci == 0;

  if (live) {
    const int R = g.fpc * th;
    for (int r = pg; r < R; r += 16) {
      const int ff = r / th, yy = r % th;
      const float* xr = s_x + cil * g.psx + (ff * sh + yy) * g.rsx;   // haloed row yy == image row yy-1
      const float* yr = s_y + col * g.psy + r * g.rsy;
      float win[CI_R][3][3];
#pragma unroll
      for (int b = 0; b < CI_R; ++b)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          win[b][ky][1] = xr[b * g.psx + ky * g.rsx + 0];
          win[b][ky][2] = xr[b * g.psx + ky * g.rsx + 1];
        }
#pragma unroll 4
      for (int xx = 0; xx < tw; ++xx) {
        float dy[CO_R];
#pragma unroll
        for (int a = 0; a < CO_R; ++a) dy[a] = yr[a * g.psy + xx];
#pragma unroll
        for (int b = 0; b < CI_R; ++b)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            win[b][ky][0] = win[b][ky][1];
            win[b][ky][1] = win[b][ky][2];
            win[b][ky][2] = xr[b * g.psx + ky * g.rsx + xx + 2];
          }
#pragma unroll
        for (int a = 0; a < CO_R; ++a) {
          bsum[a] += dy[a];
#pragma unroll
          for (int b = 0; b < CI_R; ++b)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) acc[a][b][ky * 3 + kx] = fmaf(dy[a], win[b][ky][kx], acc[a][b][ky * 3 + kx]);
        }
      }
    }
  }
  // ---- reduce over the 16 pixel groups (lanes pg = tid & 15 share a half-warp)
#pragma unroll
  for (int a = 0; a < CO_R; ++a) {
#pragma unroll
    for (int b = 0; b < CI_R; ++b)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float v = acc[a][b][t];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[a][b][t] = v;
      }
    float v = bsum[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    bsum[a] = v;
  }
  if (pg == 0 && live) {
#pragma unroll
    for (int a = 0; a < CO_R; ++a) {
      const int co = co0 + col + a;
      if (col + a >= con) continue;
#pragma unroll
      for (int b = 0; b < CI_R; ++b) {
        const int ci = ci0 + cil + b;
        if (cil + b >= cin) continue;
        float* d = p.dw + ((size_t)co * Cin + ci) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(d + t, acc[a][b][t]);
      }
      if (do_bias) atomicAdd(p.db + co, bsum[a]);
    }
  }
}

template <int CO_R, int CI_R, int NCO, int NCI>
static int launch_wgrad(const cgs_wgrad3x3_args& a, cudaStream_t st) {
  constexpr int CO_B = CO_R * NCO, CI_B = CI_R * NCI;
  WgGeom g;
  g.th = a.H < 32 ? a.H : 32;
  g.tw = a.W < 32 ? a.W : 32;
  g.tiles_y = a.H / g.th; g.tiles_x = a.W / g.tw;
  g.fpc = 1;
  // small maps: several frames per CTA so each pixel group owns >= 2 rows, while keeping >= 2 waves of CTAs
  while (g.fpc * g.th * g.tw < 1024 && (long)((a.B + 2 * g.fpc - 1) / (2 * g.fpc)) >= 148) g.fpc *= 2;
  g.rsx = (g.tw + 2) | 1; g.rsy = g.tw | 1;
  g.psx = g.fpc * (g.th + 2) * g.rsx; g.psx += (40 - (g.psx % 32)) % 32;   // plane stride == 8 (mod 32)
  g.psy = g.fpc * g.th * g.rsy;       g.psy += (40 - (g.psy % 32)) % 32;
  g.dsw = make_fastdiv(g.tw + 2); g.dsh = make_fastdiv(g.th + 2);
  g.dtw = make_fastdiv(g.tw);     g.dth = make_fastdiv(g.th);
  g.dwp = make_fastdiv(g.tw > 2 ? g.tw / 2 : 2); g.dhp = make_fastdiv(g.th > 2 ? g.th / 2 : 2);
  size_t smem = ((size_t)CI_B * g.psx + (size_t)CO_B * g.psy) * sizeof(float);
  cudaFuncSetAttribute(wgrad3x3_kernel<CO_R, CI_R, NCO, NCI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int nblk = ((a.B + g.fpc - 1) / g.fpc) * g.tiles_x * g.tiles_y;
  dim3 grid(nblk, (a.dy.C + CO_B - 1) / CO_B, (a.x.C + CI_B - 1) / CI_B);
  wgrad3x3_kernel<CO_R, CI_R, NCO, NCI><<<grid, 256, smem, st>>>(a, g);
  return check_launch("wgrad3x3");
}

}  // namespace cgs

extern "C" int cgs_conv3x3(const cgs_conv3x3_args* a, void* stream) {
  using namespace cgs;
  CGS_REQUIRE(a != nullptr, "conv3x3: null args");
  if (int e = check_src(a->src, "conv3x3.src")) return e;
  CGS_REQUIRE(a->w && a->B > 0 && a->Cout > 0, "conv3x3: bad w/B/Cout");
  CGS_REQUIRE(a->H >= 2 && a->W >= 2 && (a->H & (a->H - 1)) == 0 && (a->W & (a->W - 1)) == 0 && a->H <= 1024 && a->W <= 1024,
              "conv3x3: H,W must be powers of two >= 2 (got %dx%d)", a->H, a->W);
  CGS_REQUIRE(a->epi >= 0 && a->epi <= CGS_EPI_SPLIT_UP, "conv3x3: bad epilogue %d", a->epi);
  if (a->epi == CGS_EPI_SPLIT_UP)
    CGS_REQUIRE(a->C0 > 0 && a->C0 < a->Cout && (a->shift2 == 1 || a->shift2 == 2) && (a->out || a->out2), "conv3x3: bad SPLIT_UP args");
  else
    CGS_REQUIRE(a->out != nullptr, "conv3x3: null out");
  if (a->epi == CGS_EPI_MUL) CGS_REQUIRE(a->mul != nullptr, "conv3x3: MUL epilogue needs mul");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->precision == CGS_TF32 && conv_tc_supported(*a)) return launch_conv_tc(*a, st);
  if (a->Cout == 1) return launch_conv<1>(*a, st);
  if (a->Cout <= 4) return launch_conv<4>(*a, st);
  if (a->Cout % 16 == 0 && a->H >= 32) return launch_conv<16>(*a, st);   // small maps: more, shorter CTAs
  return launch_conv<8>(*a, st);
}

extern "C" int cgs_wgrad3x3(const cgs_wgrad3x3_args* a, void* stream) {
  using namespace cgs;
  CGS_REQUIRE(a != nullptr, "wgrad3x3: null args");
  if (int e = check_src(a->x, "wgrad3x3.x")) return e;
  if (int e = check_src(a->dy, "wgrad3x3.dy")) return e;
  CGS_REQUIRE(a->dw && a->B > 0, "wgrad3x3: null dw / B<=0");
  CGS_REQUIRE(a->H >= 2 && a->W >= 2 && (a->H & (a->H - 1)) == 0 && (a->W & (a->W - 1)) == 0 && a->H <= 1024 && a->W <= 1024,
              "wgrad3x3: H,W must be powers of two >= 2 (got %dx%d)", a->H, a->W);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->precision == CGS_TF32 && wgrad_pipe_supported(*a)) return launch_wgrad_pipe(*a, st);
  if (a->precision == CGS_TF32 && wgrad_mma_supported(*a)) return launch_wgrad_mma(*a, st);
  if (a->dy.C == 1) return launch_wgrad<1, 1, 1, 16>(*a, st);
  if (a->x.C <= 4) return launch_wgrad<2, 1, 4, 4>(*a, st);
  return launch_wgrad<2, 2, 4, 4>(*a, st);
}
