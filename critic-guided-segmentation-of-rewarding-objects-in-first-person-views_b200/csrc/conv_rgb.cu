// First encoder stage straight from the raw frames:  uint8 NHWC frames -> [/255] -> [shift_batch roll] ->
// Conv2d(3, Cout, 3, 1, 1) + bias -> ReLU -> MaxPool2d(2) -> e0 (+ arg-max bytes)      (reference main.py:185-189,
// nets.py:170-172).  Replaces frames_to_float + the generic FFMA kernel for the layer that holds 52 % of the critic's
// MACs at chfak 1.
//
// ncu on the generic kernel for this layer: 68 % issue-active but only 32 % of the executed instructions were FFMA —
// the rest was generic operand staging (runtime source modes, 64-bit index maths, 8-wide predicated channel loops for
// 3 channels) and the fp32 copy of the frames cost an extra 12.6 MB write + two 12.6 MB reads per step.  Here:
//   * persistent CTAs; raw uint8 rows (192 B per image row) arrive by 16-byte cp.async into a 3-stage ring (zfill for
//     rows outside the image): 3.4 KB per 16x64-pixel tile instead of 13.8 KB, no register staging;
//   * one small convert pass per tile (shared -> shared) applies the exact (float)u8 / 255.0f, the circular roll and the
//     zero halo, producing the planar fp32 tile the inner loop wants;
//   * the inner loop is the register-blocked direct convolution (2x2 pixels = one pool window x 8 channels per thread,
//     each loaded input feeds 72 FMAs), epilogue = bias + ReLU + first-max pool + 128-bit stores.
#include "common.cuh"

namespace cgs {

struct RgbGeom {
  int th, tiles_y, ntiles;     // tile = th rows x full width
  int raw_row, raw_chunks;     // bytes / 16-byte chunks per raw image row (W*3)
  int fpitch, fplane;          // float tile: row pitch and plane size (floats)
  int raw_stage;               // bytes per raw stage
  int off_f, off_w;            // float offsets of the fp32 tile and the weights behind the raw ring
  int rollc;                   // roll normalised to [0, W)
  int off_lut;                 // float offset of the 256-entry (float)i/255 table
  FastDiv dcol;                // divider by W + 2
};

__device__ __forceinline__ void rgb_cp16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

constexpr int RGB_NS = 3;

__global__ void __launch_bounds__(256, 3) conv_rgb_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ roll_dev,
                                                          const float* __restrict__ w, const float* __restrict__ bias, int B, int H,
                                                          int W, int Cout, float* __restrict__ e0, uint8_t* __restrict__ idx0,
                                                          const RgbGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int th = g.th, rows = th + 2;
  float* s_f = reinterpret_cast<float*>(smem_raw) + g.off_f;     // [3][rows][fpitch]
  float* s_w = reinterpret_cast<float*>(smem_raw) + g.off_w;     // [3][9][Cout]
  const uint32_t raw_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  int roll = g.rollc;
  if (roll_dev) {
    roll = *roll_dev % W;
    if (roll < 0) roll += W;
  }
  float* s_lut = reinterpret_cast<float*>(smem_raw) + g.off_lut;
  s_lut[tid] = (float)tid / 255.0f;      // exact (float)u8 / 255.0f without a division per pixel
  for (int e = tid; e < 27 * Cout; e += 256) {
    const int co = e % Cout, r = e / Cout;
    const int t = r % 9, ci = r / 9;
    s_w[e] = __ldg(w + ((size_t)co * 3 + ci) * 9 + t);
  }

  auto prefetch = [&](int tile, int s) {
    const int n = tile / g.tiles_y, y0 = (tile - n * g.tiles_y) * th;
    const uint32_t sb = raw_base + (uint32_t)(s * g.raw_stage);
    for (int c = tid; c < rows * g.raw_chunks; c += 256) {
      const int r = c / g.raw_chunks, cx = c - r * g.raw_chunks;
      const int y = y0 - 1 + r;
      const bool ok = y >= 0 && y < H;
      const uint8_t* src = ok ? frames + (size_t)(unsigned)((n * H + y) * W) * 3u + cx * 16 : frames;
      rgb_cp16(sb + (uint32_t)(r * g.raw_row + cx * 16), src, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  for (int k = 0; k < RGB_NS - 1; ++k) {
    const int t = blockIdx.x + k * gridDim.x;
    if (t < g.ntiles) prefetch(t, k); else asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  const int W2 = W >> 1, H2 = H >> 1;
  int it = 0;
  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
    const int s = it % RGB_NS;
    const int ahead = tile + (RGB_NS - 1) * gridDim.x;
    if (ahead < g.ntiles) prefetch(ahead, (it + RGB_NS - 1) % RGB_NS); else asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group %0;\n" ::"n"(RGB_NS - 1) : "memory");
    __syncthreads();      // raw rows of this tile have landed; previous tile's compute is done with s_f

    // ---- convert pass: raw u8 (pixel-major) -> planar fp32 with /255, roll and zero halo columns
    const uint8_t* raw = smem_raw + s * g.raw_stage;
    for (int e = tid; e < rows * (W + 2); e += 256) {
      const int r = fdiv(e, g.dcol), c = e - r * (W + 2);
      const int x = c - 1;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (x >= 0 && x < W) {
        int sx = x + roll;
        if (sx >= W) sx -= W;
        const uint8_t* q = raw + r * g.raw_row + sx * 3;
        v0 = s_lut[q[0]]; v1 = s_lut[q[1]]; v2 = s_lut[q[2]];
      }
      float* d = s_f + r * g.fpitch + c;
      d[0] = v0; d[g.fplane] = v1; d[2 * g.fplane] = v2;
    }
    __syncthreads();

    // ---- direct convolution: one 2x2 pool window x 8 output channels per thread iteration
    const int n = tile / g.tiles_y, y0 = (tile - n * g.tiles_y) * th;
    const int nwin = (th >> 1) * W2;
    for (int co0 = 0; co0 < Cout; co0 += 8) {
      for (int win = tid; win < nwin; win += 256) {
        const int wy = win / W2, wx = win - wy * W2;
        float acc[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float* ip = s_f + ci * g.fplane + (2 * wy) * g.fpitch + 2 * wx;
          float in[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float2 a = *reinterpret_cast<const float2*>(ip + r * g.fpitch);
            const float2 b = *reinterpret_cast<const float2*>(ip + r * g.fpitch + 2);
            in[r][0] = a.x; in[r][1] = a.y; in[r][2] = b.x; in[r][3] = b.y;
          }
          const float* wp = s_w + (size_t)ci * 9 * Cout + co0;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float4 wa = *reinterpret_cast<const float4*>(wp + t * Cout);
            const float4 wb = *reinterpret_cast<const float4*>(wp + t * Cout + 4);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            const int ky = t / 3, kx = t % 3;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[0][j] = fmaf(in[ky][kx], wv[j], acc[0][j]);
              acc[1][j] = fmaf(in[ky][kx + 1], wv[j], acc[1][j]);
              acc[2][j] = fmaf(in[ky + 1][kx], wv[j], acc[2][j]);
              acc[3][j] = fmaf(in[ky + 1][kx + 1], wv[j], acc[3][j]);
            }
          }
        }
        // bias + ReLU + max-pool (first maximum in row-major window order wins, as ATen)
        float mv[8];
        uint32_t am_lo = 0, am_hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float bb = __ldg(bias + co0 + j);
          float m = fmaxf(acc[0][j] + bb, 0.f);
          uint32_t a = 0;
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            const float v = fmaxf(acc[q][j] + bb, 0.f);
            if (v > m) { m = v; a = q; }
          }
          mv[j] = m;
          if (j < 4) am_lo |= a << (8 * j); else am_hi |= a << (8 * (j - 4));
        }
        const size_t o = (size_t)(unsigned)((n * H2 + (y0 >> 1) + wy) * W2 + wx) * (unsigned)Cout + co0;
        *reinterpret_cast<float4*>(e0 + o) = make_float4(mv[0], mv[1], mv[2], mv[3]);
        *reinterpret_cast<float4*>(e0 + o + 4) = make_float4(mv[4], mv[5], mv[6], mv[7]);
        if (idx0) *reinterpret_cast<uint2*>(idx0 + o) = make_uint2(am_lo, am_hi);
      }
    }
  }
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_conv_rgb_fwd(const uint8_t* frames, int32_t B, int32_t H, int32_t W, int32_t roll, const int32_t* roll_dev,
                                const float* w, const float* bias, int32_t Cout, float* e0, uint8_t* idx0, void* stream) {
  CGS_REQUIRE(frames && w && bias && e0 && B > 0, "conv_rgb_fwd: bad args");
  CGS_REQUIRE(H >= 16 && (H % 16) == 0 && W >= 16 && (W % 16) == 0 && W <= 512, "conv_rgb_fwd: H, W must be multiples of 16 (W <= 512)");
  CGS_REQUIRE(Cout > 0 && (Cout % 8) == 0, "conv_rgb_fwd: Cout must be a multiple of 8");
  CGS_REQUIRE(((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(e0)) & 15) == 0, "conv_rgb_fwd: unaligned buffers");
  RgbGeom g;
  g.th = 16;
  g.tiles_y = H / g.th;
  g.ntiles = B * g.tiles_y;
  g.raw_row = W * 3;                       // multiple of 48 bytes -> whole 16-byte chunks
  g.raw_chunks = g.raw_row / 16;
  g.raw_stage = (g.th + 2) * g.raw_row;
  g.fpitch = (W + 2 + 1) & ~1;             // even: float2 window loads
  g.fplane = (g.th + 2) * g.fpitch;
  g.fplane += (36 - (g.fplane % 32)) % 32; // planes 4 banks apart
  g.off_f = (RGB_NS * g.raw_stage + 15) / 16 * 4;
  g.off_w = g.off_f + 3 * g.fplane;
  g.off_w = (g.off_w + 3) & ~3;
  g.rollc = ((roll % W) + W) % W;
  g.off_lut = g.off_w + 27 * Cout;
  g.dcol = make_fastdiv(W + 2);
  const size_t smem = (size_t)(g.off_lut + 256) * sizeof(float);
  cudaFuncSetAttribute(conv_rgb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int sms = device_sms();
  CGS_REQUIRE(smem <= 160 * 1024, "conv_rgb_fwd: tile does not fit shared memory");
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm);
  int grid = sms * per_sm;
  if (grid > g.ntiles) grid = g.ntiles;
  conv_rgb_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(frames, roll_dev, w, bias, B, H, W, Cout, e0, idx0, g);
  return check_launch("conv_rgb_fwd");
}
