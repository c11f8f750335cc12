// The rest of the wide (chfak > 1) critic training step around the tcgen05 convolutions of wide_tc.cu:
//   * features.0 (3 -> C0 on 64x64, K = 27): forward + ReLU + pool + arg-max and its weight / bias gradient as bf16 mma.sync
//     kernels on the pair-duplicated frame of hg_common.cuh (one k16 MMA per filter row; the 64x64xC0 output gradient is never
//     materialised: the B fragment is selected on the fly from the pooled gradient and the arg-max bytes);
//   * the head (features.14 = a 4x4 valid convolution on a 4x4 map = a plain GEMM, crit.1, crit.4, loss, and their backward)
//     as six small TF32 GEMMs (cp.async 3-stage pipeline, mma.sync m16n8k8) + two elementwise kernels.
// Reference: nets.py:169-195 (NewCritic), main.py:185-198 (critic_pipe step).
#include <string.h>
#include "hg_common.cuh"

namespace cgs {
namespace wm {
using namespace hg;

constexpr int MAXCP = 5;                             // C0 / 8 <= 5 (chfak <= 5)
constexpr int sX = 0, sU8 = sX + PBX, sE = sU8 + 12288;                 // pair-duplicated frame | raw bytes | e0 / d e0 planes
constexpr int sI(int cp) { return sE + cp * 16384; }                    // arg-max bytes [cp][32][32][8]
constexpr int sWf(int cp) { return sI(cp) + cp * 8192; }                // forward: weight fragments [3][cp][32] uint2 | bias
constexpr int conv0_smem(int cp) { return sWf(cp) + 3 * cp * 256 + cp * 32 + 64; }

__device__ __forceinline__ void stage_frame(const uint8_t* __restrict__ u8, uint8_t* __restrict__ dst, int roll, float c3, int tid) {
  for (int e = tid; e < 4096; e += NT) {
    const int y = e >> 6, x = e & 63;
    const uint8_t* s = u8 + (y * 64 + ((x + roll) & 63)) * 3;
    constexpr float k = 1.f / 255.f;                 // bf16(b * fl(1/255)) == bf16(b / 255.0f) for all 256 bytes
    const uint2 q = make_uint2(pack_bf16(__fmul_rn((float)s[0], k), __fmul_rn((float)s[1], k)), pack_bf16(__fmul_rn((float)s[2], k), c3));
    uint8_t* row = dst + (size_t)(y + 1) * (PX * 16);
    *reinterpret_cast<uint2*>(row + (x + 1) * 16) = q;
    *reinterpret_cast<uint2*>(row + x * 16 + 8) = q;
  }
}

// ---- features.0 forward: uint8 frame -> /255 + roll -> conv + ReLU + pool + arg-max -> e0 [B][CP][32][32][8] bf16, idx0 uint8
__global__ void __launch_bounds__(NT, 1) wide_conv0_fwd_kernel(const uint8_t* __restrict__ frames, int B, int roll, const int* __restrict__ roll_dev,
                                                               const float* __restrict__ w0, const float* __restrict__ b0, int CP,
                                                               __nv_bfloat16* __restrict__ e0, uint8_t* __restrict__ idx0) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, odd = g & 1;
  const int lj = lane >> 3, lr = lane & 7, pixoff = 2 * lr + (lj & 1), chunk = lj >> 1;   // A row r -> pixel 2 (r & 7) + (r >> 3)
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(smraw);
  const int oI = sE + CP * 16384, oW = oI + CP * 8192, oB = oW + 3 * CP * 256;
  uint2* sW = reinterpret_cast<uint2*>(smraw + oW);
  float* sB = reinterpret_cast<float*>(smraw + oB);
  pdl_trigger();
  pdl_wait();
  if ((int)blockIdx.x < B) {
    const uint8_t* src = frames + (size_t)blockIdx.x * 12288;
    for (int c = tid; c < 768; c += NT) cp_async16(smb + sU8 + c * 16, src + c * 16);
    cp_async_commit();
  }
  for (int e = tid; e < PBX / 16; e += NT) reinterpret_cast<uint4*>(smraw + sX)[e] = make_uint4(0u, 0u, 0u, 0u);
  // B fragments of step ky, channel tile nt: k = kx * 4 + c (pair-duplicated frame), n = output channel nt * 8 + g
  for (int e = tid; e < 3 * CP * 32; e += NT) {
    const int ln = e & 31, nt = (e >> 5) % CP, ky = (e >> 5) / CP, gg = ln >> 2, tt = ln & 3, co = nt * 8 + gg;
    auto wk = [&](int k) {
      const int kx = k >> 2, c = k & 3;
      return (kx < 3 && c < 3) ? __ldg(w0 + ((co * 3 + c) * 3 + ky) * 3 + kx) : 0.f;
    };
    sW[e] = make_uint2(pack_bf16(wk(2 * tt), wk(2 * tt + 1)), pack_bf16(wk(2 * tt + 8), wk(2 * tt + 9)));
  }
  for (int e = tid; e < CP * 8; e += NT) sB[e] = __ldg(b0 + e);
  int rl = roll_dev ? *roll_dev : roll;
  rl = ((rl % 64) + 64) & 63;
  for (int n = blockIdx.x; n < B; n += gridDim.x) {
    cp_async_wait_all();
    __syncthreads();                                 // frame bytes landed; the previous frame's planes have been copied out
    stage_frame(smraw + sU8, smraw + sX, rl, 0.f, tid);
    __syncthreads();
    if (n + (int)gridDim.x < B) {
      const uint8_t* src = frames + (size_t)(n + gridDim.x) * 12288;
      for (int c = tid; c < 768; c += NT) cp_async16(smb + sU8 + c * 16, src + c * 16);
      cp_async_commit();
    }
    const int x0 = (warp & 3) * 16, r0 = (warp >> 2) * 16;
    const uint32_t aA = smb + sX + (uint32_t)((r0 * PX + x0 + pixoff + 2 * chunk) * 16);
    for (int nt = 0; nt < CP; ++nt) {
      uint2 w[3][1][1];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) w[ky][0][0] = sW[(ky * CP + nt) * 32 + lane];
      const float bias0 = sB[nt * 8 + 2 * t], bias1 = sB[nt * 8 + 2 * t + 1];
      // the A rows of a fragment are permuted (row g = pixel 2g, row g + 8 = pixel 2g + 1 of the 16-pixel strip), so a thread holds
      // the x pair of both rows: the whole 2x2 window of pooled pixel (x0 / 2 + g), channels 2t and 2t + 1, no shuffles
      uint32_t* dE = reinterpret_cast<uint32_t*>(smraw + sE) + ((nt * 32 + (r0 >> 1)) * 32 + (x0 >> 1) + g) * 4 + t;
      unsigned short* dI = reinterpret_cast<unsigned short*>(smraw + oI) + ((nt * 32 + (r0 >> 1)) * 32 + (x0 >> 1) + g) * 4 + t;
      slide_bf<16, 1, 1>(
          w, [&](int i, uint32_t(&a)[1][4]) { ldsm4(a[0], aA + (uint32_t)(i * (PX * 16))); },
          [&](int e, int, const float(&top)[4], const float(&bot)[4]) {
            float v[2];
            uint32_t id[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float bb = j ? bias1 : bias0;
              const float p0 = top[j] + bb, p1 = top[2 + j] + bb, p2 = bot[j] + bb, p3 = bot[2 + j] + bb;
              const float m01 = fmaxf(p0, p1), m23 = fmaxf(p2, p3);
              const uint32_t i01 = p1 > p0 ? 1u : 0u, i23 = p3 > p2 ? 3u : 2u;
              float m = fmaxf(m01, m23);
              uint32_t ix = m23 > m01 ? i23 : i01;
              if (!(m > 0.f)) { m = 0.f; ix = 4u; }
              v[j] = m; id[j] = ix;
            }
            dE[(e >> 1) * 128] = pack_bf16(v[0], v[1]);
            dI[(e >> 1) * 128] = (unsigned short)(id[0] | (id[1] << 8));
          });
    }
    __syncthreads();
    {   // planes out: the chunk-planar layout of a frame is contiguous in HBM
      const uint4* s = reinterpret_cast<const uint4*>(smraw + sE);
      uint4* d = reinterpret_cast<uint4*>(e0 + (size_t)n * CP * 8192);
      for (int e = tid; e < CP * 1024; e += NT) d[e] = s[e];
      const uint4* si = reinterpret_cast<const uint4*>(smraw + oI);
      uint4* di = reinterpret_cast<uint4*>(idx0 + (size_t)n * CP * 8192);
      for (int e = tid; e < CP * 512; e += NT) di[e] = si[e];
    }
  }
  cp_async_wait_all();
}

// ---- features.0 weight + bias gradient: partial[cta][ky][m = kx * 4 + c][co], c = 3 is the ones channel (bias at the centre tap)
__global__ void __launch_bounds__(NT, 1) wide_conv0_wgrad_kernel(const uint8_t* __restrict__ frames, int B, int roll, const int* __restrict__ roll_dev,
                                                                 const __nv_bfloat16* __restrict__ de0, const uint8_t* __restrict__ idx0, int CP,
                                                                 float* __restrict__ partials) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int lj = lane >> 3, lr = lane & 7, tsel = lj & 1, tpix = lr + 8 * (lj >> 1);
  const uint32_t smb = (uint32_t)__cvta_generic_to_shared(smraw);
  const int oI = sE + CP * 16384;
  const unsigned short* uDE = reinterpret_cast<const unsigned short*>(smraw + sE);
  const uint8_t* bI = smraw + oI;
  pdl_trigger();
  for (int e = tid; e < PBX / 16; e += NT) reinterpret_cast<uint4*>(smraw + sX)[e] = make_uint4(0u, 0u, 0u, 0u);
  pdl_wait();
  int rl = roll_dev ? *roll_dev : roll;
  rl = ((rl % 64) + 64) & 63;
  float acc[MAXCP][3][4];
#pragma unroll
  for (int i = 0; i < MAXCP * 12; ++i) acc[i / 12][(i / 4) % 3][i & 3] = 0.f;
  for (int n = blockIdx.x; n < B; n += gridDim.x) {
    __syncthreads();                                 // the previous frame's operands are no longer read
    {
      const uint8_t* src = frames + (size_t)n * 12288;
      for (int c = tid; c < 768; c += NT) cp_async16(smb + sU8 + c * 16, src + c * 16);
      const uint8_t* sd = reinterpret_cast<const uint8_t*>(de0 + (size_t)n * CP * 8192);
      for (int c = tid; c < CP * 1024; c += NT) cp_async16(smb + sE + c * 16, sd + c * 16);
      const uint8_t* si = idx0 + (size_t)n * CP * 8192;
      for (int c = tid; c < CP * 512; c += NT) cp_async16(smb + oI + c * 16, si + c * 16);
      cp_async_commit();
      cp_async_wait_all();
    }
    __syncthreads();
    stage_frame(smraw + sU8, smraw + sX, rl, 1.f, tid);
    __syncthreads();
    const int x0 = (warp & 3) * 16, y0 = (warp >> 2) * 16;
    uint32_t bq[MAXCP][3][2];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      uint32_t a[4];
      ldsm4t(a, smb + sX + (uint32_t)(((y0 + i) * PX + x0 + tpix + 2 * tsel) * 16));
      if (i < 16) {
        const int y = y0 + i, py = y >> 1;
#pragma unroll
        for (int nt = 0; nt < MAXCP; ++nt)
          if (nt < CP) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {           // b0: pixels x0 + 2t, 2t+1 -> pooled px (x0 >> 1) + t; b1: + 8 -> pooled px + 4
              const int o = ((nt * 32 + py) * 32 + (x0 >> 1) + t + 4 * hh) * 8 + g;
              const uint32_t v = uDE[o], idx = bI[o];
              const uint32_t lo = idx == (uint32_t)((y & 1) << 1) ? v : 0u, hi = idx == (uint32_t)(((y & 1) << 1) | 1) ? v : 0u;
              bq[nt][i % 3][hh] = lo | (hi << 16);
            }
          }
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = i - ky;
        if (yy >= 0 && yy < 16) {
#pragma unroll
          for (int nt = 0; nt < MAXCP; ++nt)
            if (nt < CP) mma_bf16(acc[nt][ky], a, bq[nt][yy % 3][0], bq[nt][yy % 3][1]);
        }
      }
    }
  }
  __syncthreads();
  // 16 warps -> one partial per CTA (fixed order): scratch [16 warps][CP][3][128] in the frame / plane region
  float* scr = reinterpret_cast<float*>(smraw);
#pragma unroll
  for (int nt = 0; nt < MAXCP; ++nt)
    if (nt < CP) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
        *reinterpret_cast<float4*>(scr + ((warp * CP + nt) * 3 + ky) * 128 + lane * 4) = make_float4(acc[nt][ky][0], acc[nt][ky][1], acc[nt][ky][2], acc[nt][ky][3]);
    }
  __syncthreads();
  // fragment value q of lane (g, t): row m = g + 8 (q >> 1), column 2t + (q & 1)
  for (int e = tid; e < CP * 3 * 128; e += NT) {
    const int nt = e / 384, ky = (e >> 7) % 3, ln = (e >> 2) & 31, q = e & 3;
    const int m = (ln >> 2) + 8 * (q >> 1), co = nt * 8 + 2 * (ln & 3) + (q & 1);
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) v += scr[((w * CP + nt) * 3 + ky) * 128 + ln * 4 + q];
    partials[(((size_t)blockIdx.x * 3 + ky) * 16 + m) * (CP * 8) + co] = v;
  }
}

__global__ void wide_conv0_reduce_kernel(const float* __restrict__ part, int ncta, int C0, float* __restrict__ dw, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;      // one warp per element, lanes over the CTAs
  if (e >= 3 * 16 * C0) return;
  const int co = e % C0, m = (e / C0) & 15, ky = e / (C0 * 16), kx = m >> 2, c = m & 3;
  const bool is_w = kx < 3 && c < 3, is_b = (ky == 1 && kx == 1 && c == 3);
  if (!is_w && !is_b) return;
  float s = 0.f;
  for (int i = lane; i < ncta; i += 32) s += part[(size_t)i * 48 * C0 + e];
  s = warp_sum(s);
  if (lane == 0) {
    if (is_w) dw[((co * 3 + c) * 3 + ky) * 3 + kx] += s;
    else db[co] += s;
  }
}

// --------------------------------------------------------------------------------------------------------------------
// C[M][N] (+)= op(A)[M][K] * op(B)[K][N], fp32 in HBM, TF32 tensor-core operands, fp32 accumulation.
//   AKC: A[m * lda + k] else A[k * lda + m];  BKC: B[n * ldb + k] else B[k * ldb + n]  (the contiguous extents are multiples of 4)
// epilogue: + bias[n], * (gate[m][n] > 0), ReLU, accumulate into C
struct GemmP {
  const float *A, *B, *bias, *gate;
  float* C;
  int M, N, K, lda, ldb, ldc, relu, accumulate;
  int splits;                      // gridDim.z: K is cut into `splits` runs of k-tiles; partial tiles meet in ws, the last CTA of a tile sums them
  float* ws;                       // [splits][M][N]
  int* counters;                   // one per (tile x, tile y), zero between launches (the last CTA resets its own)
};
constexpr int GM = 64, GN = 32, GK = 32, GST = 6;     // 6 stages: these GEMMs are short K loops bound by L2 latency, not bytes
constexpr int GA_FLOATS = 2304, GB_FLOATS = 1280;    // [64][36] or [32][72]; [32][36] or [32][40]

__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <bool AKC, bool BKC>
__global__ void __launch_bounds__(128) wide_gemm_kernel(const GemmP p) {
  extern __shared__ __align__(16) float gsm[];
  float(*sA)[GA_FLOATS] = reinterpret_cast<float(*)[GA_FLOATS]>(gsm);
  float(*sB)[GB_FLOATS] = reinterpret_cast<float(*)[GB_FLOATS]>(gsm + GST * GA_FLOATS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  pdl_trigger();
  pdl_wait();
  const int nk_all = (p.K + GK - 1) / GK, per = (nk_all + p.splits - 1) / p.splits;
  const int kt0 = blockIdx.z * per, nk = max(0, min(per, nk_all - kt0));
  // per-thread chunk descriptors, fixed over the K loop (a single warp per scheduler runs these loops: every instruction counts):
  // A: 4 chunks of 16 bytes per thread and k-tile, B: 2
  const float* pa[4];
  const float* pb[2];
  uint32_t da[4], db[2];
  int ka[4], kb[2];                // k offset of the chunk inside a k-tile (-1: the chunk is outside M / N for good)
  const uint32_t sa0 = (uint32_t)__cvta_generic_to_shared(gsm), sb0 = sa0 + GST * GA_FLOATS * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tid + 128 * j;
    if (AKC) {
      const int r = c >> 3, kc = (c & 7) * 4;
      pa[j] = p.A + (size_t)min(m0 + r, p.M - 1) * p.lda + kc; da[j] = (r * 36 + kc) * 4; ka[j] = (m0 + r) < p.M ? kc : -1;
    } else {
      const int r = c >> 4, mc = (c & 15) * 4;
      pa[j] = p.A + (size_t)r * p.lda + min(m0 + mc, p.M - 4); da[j] = (r * 72 + mc) * 4; ka[j] = (m0 + mc) < p.M ? r : -1;
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = tid + 128 * j;
    if (BKC) {
      const int r = c >> 3, kc = (c & 7) * 4;
      pb[j] = p.B + (size_t)min(n0 + r, p.N - 1) * p.ldb + kc; db[j] = (r * 36 + kc) * 4; kb[j] = (n0 + r) < p.N ? kc : -1;
    } else {
      const int r = c >> 3, nc = (c & 7) * 4;
      pb[j] = p.B + (size_t)r * p.ldb + min(n0 + nc, p.N - 4); db[j] = (r * 40 + nc) * 4; kb[j] = (n0 + nc) < p.N ? r : -1;
    }
  }
  const size_t astep = AKC ? (size_t)1 : (size_t)p.lda, bstep = BKC ? (size_t)1 : (size_t)p.ldb;
  auto load = [&](int kt, int st) {
    const int k0 = (kt0 + kt) * GK;
    const uint32_t a = sa0 + st * (GA_FLOATS * 4), b = sb0 + st * (GB_FLOATS * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = ka[j] >= 0 && (k0 + ka[j]) < p.K;
      cp16z(a + da[j], ok ? pa[j] + (size_t)k0 * astep : p.A, ok);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const bool ok = kb[j] >= 0 && (k0 + kb[j]) < p.K;
      cp16z(b + db[j], ok ? pb[j] + (size_t)k0 * bstep : p.B, ok);
    }
    cp_async_commit();
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i >> 2][i & 3] = 0.f;
  for (int s = 0; s < GST - 1; ++s) {
    if (s < nk) load(s, s);
    else cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(GST - 2));
    __syncthreads();
    if (kt + GST - 1 < nk) load(kt + GST - 1, (kt + GST - 1) % GST);
    else cp_async_commit();
    const uint32_t* a = reinterpret_cast<const uint32_t*>(sA[kt % GST]);
    const uint32_t* b = reinterpret_cast<const uint32_t*>(sB[kt % GST]);
    const int mr = warp * 16 + g;
#pragma unroll
    for (int k8 = 0; k8 < 4; ++k8) {
      const int k = k8 * 8 + t;
      uint32_t a0, a1, a2, a3;
      if (AKC) { a0 = a[mr * 36 + k]; a1 = a[(mr + 8) * 36 + k]; a2 = a[mr * 36 + k + 4]; a3 = a[(mr + 8) * 36 + k + 4]; }
      else { a0 = a[k * 72 + mr]; a1 = a[k * 72 + mr + 8]; a2 = a[(k + 4) * 72 + mr]; a3 = a[(k + 4) * 72 + mr + 8]; }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int nn = nt * 8 + g;
        uint32_t b0, b1;
        if (BKC) { b0 = b[nn * 36 + k]; b1 = b[nn * 36 + k + 4]; }
        else { b0 = b[k * 40 + nn]; b1 = b[(k + 4) * 40 + nn]; }
        mma_tf32(acc[nt], a0, a1, a2, a3, b0, b1);
      }
    }
  }
  cp_async_wait_all();
  if (p.splits > 1) {
    // partial tile -> ws[z]; the last CTA to arrive for this (x, y) tile sums the `splits` partials in order z = 0, 1, ...
    float* mine = p.ws + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m0 + warp * 16 + g + 8 * (q >> 1), n = n0 + nt * 8 + 2 * t + (q & 1);
        if (m < p.M && n < p.N) __stcg(mine + (size_t)m * p.N + n, acc[nt][q]);
      }
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      int* ctr = p.counters + blockIdx.y * gridDim.x + blockIdx.x;
      const int prev = atomicAdd(ctr, 1);
      s_last = (prev == p.splits - 1);
      if (s_last) *ctr = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m0 + warp * 16 + g + 8 * (q >> 1), n = n0 + nt * 8 + 2 * t + (q & 1);
        float pz[8];
#pragma unroll
        for (int z = 0; z < 8; ++z) pz[z] = (z < p.splits && m < p.M && n < p.N) ? __ldcg(p.ws + ((size_t)z * p.M + m) * p.N + n) : 0.f;
        float v = 0.f;
#pragma unroll
        for (int z = 0; z < 8; ++z) v += pz[z];
        acc[nt][q] = v;
      }
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = m0 + warp * 16 + g + 8 * (q >> 1), n = n0 + nt * 8 + 2 * t + (q & 1);
      if (m < p.M && n < p.N) {
        float v = acc[nt][q];
        if (p.bias) v += __ldg(p.bias + n);
        if (p.gate && !(__ldg(p.gate + (size_t)m * p.ldc + n) > 0.f)) v = 0.f;
        if (p.relu) v = fmaxf(v, 0.f);
        float* c = p.C + (size_t)m * p.ldc + n;
        *c = p.accumulate ? *c + v : v;
      }
    }
}

// ---- crit.3 Dropout, crit.4 Linear(nb, 1), Sigmoid, MSE / BCE per frame (one warp per frame): pred, the frame's loss term, dz = d loss
// / d logit, dV = d loss / d(crit.1 output) (ReLU and dropout applied) and U = dz * V * mask (whose column sum is crit.4's gradient)
__global__ void __launch_bounds__(256) wide_head_mid_kernel(const float* __restrict__ V, const float* __restrict__ mv, const float* __restrict__ wl2,
                                                            const float* __restrict__ bl2, const float* __restrict__ target, int B, int nb,
                                                            float gscale, int bce, float* __restrict__ pred, float* __restrict__ lterm,
                                                            float* __restrict__ dV, float* __restrict__ dz_out, float* __restrict__ U) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float s = 0.f;
  for (int j = lane; j < nb; j += 32) s += __ldg(wl2 + j) * __ldg(V + (size_t)b * nb + j) * (mv ? __ldg(mv + (size_t)b * nb + j) : 1.f);
  const float z = warp_sum(s) + __ldg(bl2), pr = sigmoidf_(z), y = __ldg(target + b);
  float dl, lt;
  if (bce) {
    lt = -(y * fmaxf(logf(pr), -100.f) + (1.f - y) * fmaxf(logf(1.f - pr), -100.f));
    dl = gscale * (pr - y) / fmaxf(pr * (1.f - pr), 1e-12f) * pr * (1.f - pr);
  } else {
    lt = (pr - y) * (pr - y);
    dl = gscale * 2.f * (pr - y) * pr * (1.f - pr);
  }
  if (lane == 0) { pred[b] = pr; dz_out[b] = dl; lterm[b] = lt; }
  for (int j = lane; j < nb; j += 32) {
    const float v = __ldg(V + (size_t)b * nb + j), m = mv ? __ldg(mv + (size_t)b * nb + j) : 1.f;
    dV[(size_t)b * nb + j] = v > 0.f ? dl * __ldg(wl2 + j) * m : 0.f;
    U[(size_t)b * nb + j] = dl * v * m;
  }
}

// up to 5 column sums in one launch (blockIdx.y = job): out[j] (+)= scale * sum_b X[b][j].  Block (32 columns x 8 row groups): the
// row groups run independent loads, then meet in shared memory in fixed order.
struct ColJob {
  const float* X;
  float* out;
  int n;
  float scale;
  int accumulate;
};
struct ColJobs {
  ColJob j[5];
};
__global__ void __launch_bounds__(256) wide_colsums_kernel(const ColJobs jobs, int B) {
  __shared__ float sm[8][33];
  pdl_trigger();
  pdl_wait();
  const ColJob jb = jobs.j[blockIdx.y];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5, col = blockIdx.x * 32 + c;
  if (blockIdx.x * 32 >= jb.n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (col < jb.n) {
    int b = r;
    for (; b + 24 < B; b += 32) {
      s0 += __ldg(jb.X + (size_t)b * jb.n + col); s1 += __ldg(jb.X + (size_t)(b + 8) * jb.n + col);
      s2 += __ldg(jb.X + (size_t)(b + 16) * jb.n + col); s3 += __ldg(jb.X + (size_t)(b + 24) * jb.n + col);
    }
    for (; b < B; b += 8) s0 += __ldg(jb.X + (size_t)b * jb.n + col);
  }
  sm[r][c] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (r == 0 && col < jb.n) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][c];
    s *= jb.scale;
    jb.out[col] = jb.accumulate ? jb.out[col] + s : s;
  }
}

// d e3 [B][C3][4][4] fp32 -> Dropout + MaxPool + ReLU backward -> dY3 [B][C3/8][8][8][8] bf16 (features.10's output gradient)
__global__ void wide_unpool3_kernel(const float* __restrict__ de3, const uint8_t* __restrict__ idx3, const float* __restrict__ m3, int B, int C3,
                                    __nv_bfloat16* __restrict__ dy3) {
  pdl_trigger();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int CP = C3 >> 3;
  if (e >= B * CP * 64) return;
  const int x = e & 7, y = (e >> 3) & 7, cg = (e >> 6) % CP, n = e / (64 * CP);
  const int py = y >> 1, px = x >> 1, pos = ((y & 1) << 1) | (x & 1);
  const uint2 ib = __ldg(reinterpret_cast<const uint2*>(idx3) + ((size_t)(n * CP + cg) * 4 + py) * 4 + px);
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t id = ((c < 4 ? ib.x : ib.y) >> (8 * (c & 3))) & 0xffu;
    float d = 0.f;
    if ((int)id == pos) {
      d = __ldg(de3 + ((size_t)(n * C3 + cg * 8 + c) * 4 + py) * 4 + px);
      if (m3) d *= __ldg(m3 + ((size_t)(n * 4 + py) * 4 + px) * C3 + cg * 8 + c);
    }
    v[c] = d;
  }
  reinterpret_cast<uint4*>(dy3)[e] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

}  // namespace wm
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_wide_conv0_fwd(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const float* w0, const float* b0,
                                  int32_t C0, void* e0, uint8_t* idx0, void* stream) {
  CGS_REQUIRE(frames && w0 && b0 && e0 && idx0 && B > 0, "wide_conv0_fwd: bad args");
  CGS_REQUIRE((C0 % 8) == 0 && C0 >= 8 && C0 <= 8 * wm::MAXCP, "wide_conv0_fwd: C0 = %d unsupported (multiple of 8, <= %d)", C0, 8 * wm::MAXCP);
  CGS_REQUIRE(((uintptr_t)frames & 15) == 0, "wide_conv0_fwd: frames must be 16-byte aligned");
  const int CP = C0 / 8, smem = wm::conv0_smem(CP);
  cudaFuncSetAttribute(wm::wide_conv0_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int grid = device_sms();
  if (grid > B) grid = B;
  launch_pdl(wm::wide_conv0_fwd_kernel, dim3(grid), dim3(hg::NT), (size_t)smem, (cudaStream_t)stream, frames, (int)B, (int)roll, (const int*)roll_dev, w0, b0, CP,
             (__nv_bfloat16*)e0, idx0);
  return check_launch("wide_conv0_fwd");
}

extern "C" int cgs_wide_conv0_wgrad(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const void* de0, const uint8_t* idx0,
                                    int32_t C0, float* dw0, float* db0, float* workspace, int64_t workspace_floats, void* stream) {
  CGS_REQUIRE(frames && de0 && idx0 && dw0 && db0 && workspace && B > 0, "wide_conv0_wgrad: bad args");
  CGS_REQUIRE((C0 % 8) == 0 && C0 >= 8 && C0 <= 8 * wm::MAXCP, "wide_conv0_wgrad: C0 = %d unsupported", C0);
  const int CP = C0 / 8, smem = wm::conv0_smem(CP);
  int grid = device_sms();
  if (grid > B) grid = B;
  CGS_REQUIRE(workspace_floats >= (int64_t)grid * 48 * C0, "wide_conv0_wgrad: workspace too small");
  cudaFuncSetAttribute(wm::wide_conv0_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  launch_pdl(wm::wide_conv0_wgrad_kernel, dim3(grid), dim3(hg::NT), (size_t)smem, (cudaStream_t)stream, frames, (int)B, (int)roll, (const int*)roll_dev,
             (const __nv_bfloat16*)de0, idx0, CP, workspace);
  int rc = check_launch("wide_conv0_wgrad");
  if (rc) return rc;
  const int n = 48 * C0;
  launch_pdl(wm::wide_conv0_reduce_kernel, dim3((n + 7) / 8), dim3(256), 0, (cudaStream_t)stream, (const float*)workspace, grid, (int)C0, dw0, db0);
  return check_launch("wide_conv0_wgrad.reduce");
}

extern "C" int cgs_wide_gemm(const float* A, int32_t a_k_contiguous, int32_t lda, const float* Bm, int32_t b_k_contiguous, int32_t ldb, float* Cm,
                             int32_t ldc, int32_t M, int32_t N, int32_t K, const float* bias, const float* gate, int32_t relu, int32_t accumulate,
                             int32_t splits, float* ws, int32_t* counters, void* stream) {
  CGS_REQUIRE(A && Bm && Cm && M > 0 && N > 0 && K > 0, "wide_gemm: bad args");
  CGS_REQUIRE((lda % 4) == 0 && (ldb % 4) == 0 && (((uintptr_t)A | (uintptr_t)Bm) & 15) == 0, "wide_gemm: operands must be 16-byte aligned rows");
  CGS_REQUIRE(((a_k_contiguous ? K : M) % 4) == 0 && ((b_k_contiguous ? K : N) % 4) == 0, "wide_gemm: contiguous extents must be multiples of 4");
  const dim3 grid((N + wm::GN - 1) / wm::GN, (M + wm::GM - 1) / wm::GM, splits > 1 ? splits : 1);
  CGS_REQUIRE(splits <= 1 || (ws && counters && grid.x * grid.y <= 4096 && splits <= 8),
              "wide_gemm: split-K (<= 8) needs a workspace [splits][M][N] and <= 4096 counters");
  CGS_REQUIRE((a_k_contiguous || M >= 4) && (b_k_contiguous || N >= 4), "wide_gemm: M, N >= 4");
  wm::GemmP p{A, Bm, bias, gate, Cm, M, N, K, lda, ldb, ldc, relu, accumulate, (int)grid.z, ws, counters};
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = wm::GST * (wm::GA_FLOATS + wm::GB_FLOATS) * 4;
  if (a_k_contiguous && b_k_contiguous) {
    cudaFuncSetAttribute(wm::wide_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_pdl(wm::wide_gemm_kernel<true, true>, grid, dim3(128), (size_t)smem, st, p);
  } else if (a_k_contiguous) {
    cudaFuncSetAttribute(wm::wide_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_pdl(wm::wide_gemm_kernel<true, false>, grid, dim3(128), (size_t)smem, st, p);
  } else if (!b_k_contiguous) {
    cudaFuncSetAttribute(wm::wide_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_pdl(wm::wide_gemm_kernel<false, false>, grid, dim3(128), (size_t)smem, st, p);
  }
  else CGS_REQUIRE(false, "wide_gemm: A MN-contiguous with B K-contiguous is not built");
  return check_launch("wide_gemm");
}

extern "C" int cgs_wide_head_mid(const float* V, const float* mv, const float* wl2, const float* bl2, const float* target, int32_t B, int32_t nb,
                                 float loss_grad, int32_t bce, float* pred, float* lterm, float* dV, float* dz, float* U, void* stream) {
  CGS_REQUIRE(V && wl2 && bl2 && target && pred && lterm && dV && dz && U && B > 0 && nb > 0, "wide_head_mid: bad args");
  launch_pdl(wm::wide_head_mid_kernel, dim3((B + 7) / 8), dim3(256), 0, (cudaStream_t)stream, V, mv, wl2, bl2, target, (int)B, (int)nb,
             loss_grad / (float)B, (int)bce, pred, lterm, dV, dz, U);
  return check_launch("wide_head_mid");
}

extern "C" int cgs_wide_colsums(const cgs_wide_coljob* jobs, int32_t njobs, int32_t B, void* stream) {
  CGS_REQUIRE(jobs && njobs >= 1 && njobs <= 5 && B > 0, "wide_colsums: 1..5 jobs");
  wm::ColJobs js;
  int maxn = 0;
  for (int i = 0; i < njobs; ++i) {
    CGS_REQUIRE(jobs[i].X && jobs[i].out && jobs[i].n > 0, "wide_colsums: bad job %d", i);
    js.j[i] = wm::ColJob{jobs[i].X, jobs[i].out, jobs[i].n, jobs[i].scale, jobs[i].accumulate};
    if (jobs[i].n > maxn) maxn = jobs[i].n;
  }
  launch_pdl(wm::wide_colsums_kernel, dim3((maxn + 31) / 32, njobs), dim3(256), 0, (cudaStream_t)stream, js, (int)B);
  return check_launch("wide_colsums");
}

extern "C" int cgs_wide_unpool3(const float* de3, const uint8_t* idx3, const float* m3, int32_t B, int32_t C3, void* dy3, void* stream) {
  CGS_REQUIRE(de3 && idx3 && dy3 && B > 0 && (C3 % 8) == 0, "wide_unpool3: bad args");
  const int n = B * (C3 / 8) * 64;
  launch_pdl(wm::wide_unpool3_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, de3, idx3, m3, (int)B, (int)C3, (__nv_bfloat16*)dy3);
  return check_launch("wide_unpool3");
}
