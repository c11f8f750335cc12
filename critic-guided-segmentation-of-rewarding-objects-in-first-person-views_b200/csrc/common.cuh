// Shared device/host helpers for the cgs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "cgs_b200.h"

namespace cgs {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
// SM count of the CURRENT device (cached per device: one process may drive several GPUs).  Function attributes
// (cudaFuncAttributeMaxDynamicSharedMemorySize) are per device too, so launchers set them on every call: it is a host-side
// table write, not a stream operation.
int device_sms();

#define CGS_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      cgs::set_error(__VA_ARGS__);        \
      return CGS_EINVAL;                  \
    }                                     \
  } while (0)

constexpr float kLeakySlope = 0.01f;  // nets.py:462

// Value of conv-operand element (n, y, x, c) at resolution H x W.  In-bounds only.
__device__ __forceinline__ float src_load(const cgs_src& s, int n, int y, int x, int c, int H, int W) {
  switch (s.mode) {
    case CGS_SRC_PLAIN: {
      size_t o = (((size_t)n * H + y) * W + x) * s.C + c;
      float v = __ldg(s.a + o);
      if (s.b) v *= __ldg(s.b + o);
      return v;
    }
    case CGS_SRC_CATUP: {
      if (c < s.C0) return __ldg(s.a + (((size_t)n * H + y) * W + x) * s.C0 + c);
      const int C1 = s.C - s.C0, h2 = H >> s.shift, w2 = W >> s.shift;
      return __ldg(s.b + (((size_t)n * h2 + (y >> s.shift)) * w2 + (x >> s.shift)) * C1 + (c - s.C0));
    }
    case CGS_SRC_POOLBWD: {
      const int h2 = H >> 1, w2 = W >> 1;
      size_t o = (((size_t)n * h2 + (y >> 1)) * w2 + (x >> 1)) * s.C + c;
      const int pos = ((y & 1) << 1) | (x & 1);
      // ReLU'(pre-act) == (pooled output > 0) at the arg-max position; elsewhere no gradient.
      if (s.idx[o] != pos || !(__ldg(s.b + o) > 0.f)) return 0.f;
      return __ldg(s.a + o);
    }
    case CGS_SRC_SIGGRAD: {
      size_t o = (((size_t)n * H + y) * W + x) * s.C + c;
      float z = __ldg(s.b + o);
      return __ldg(s.a + o) * z * (1.f - z);
    }
    case CGS_SRC_LEAKYGRAD: {
      size_t o = (((size_t)n * H + y) * W + x) * s.C + c;
      float g = __ldg(s.a + o);
      return __ldg(s.b + o) > 0.f ? g : g * kLeakySlope;
    }
  }
  return 0.f;
}

// ---- division-free index maths: q = n / d via one UMULHI with M = ceil(2^32 / d), exact for n*d < 2^32
struct FastDiv {
  uint32_t m;
  int d;
};
inline FastDiv make_fastdiv(int d) { return FastDiv{(uint32_t)(0xFFFFFFFFu / (uint32_t)d + 1u), d}; }   // d >= 2
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return (int)__umulhi((uint32_t)n, f.m); }

// Tile geometry shared by the wgrad kernels (conv_fp32.cu, wgrad_mma.cu).
struct WgGeom {
  int th, tw, fpc, tiles_y, tiles_x;
  int rsx, psx, rsy, psy;
  FastDiv dsw, dsh, dtw, dth;
  FastDiv dwp, dhp;   // dividers by tw/2, th/2 (pooled-granularity staging)
  int xplanes;        // X planes actually allocated (min(8, Cin)): the RGB layer needs 3, not 8
  int nblk;           // pixel tiles (x frame groups) in total; CTAs are persistent over them
};

// Up to 8 consecutive floats (channels) of one pixel -> registers; two 128-bit loads when aligned.
__device__ __forceinline__ void load8(const float* __restrict__ p, int n, bool vec, float (&v)[8]) {
  if (vec && n == 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = i < n ? __ldg(p + i) : 0.f;
    if (n > 4) {
#pragma unroll
      for (int i = 4; i < 8; ++i) v[i] = i < n ? __ldg(p + i) : 0.f;
    } else {
      v[4] = v[5] = v[6] = v[7] = 0.f;
    }
  }
}

// Channels [c0, c0+cn) (cn <= 8) of conv-operand pixel (n, y, x), in-bounds, into v[0..cn).
__device__ __forceinline__ void src_load8(const cgs_src& s, int n, int y, int x, int c0, int cn, int H, int W, float (&v)[8]) {
  switch (s.mode) {
    case CGS_SRC_PLAIN: {
      const size_t o = (size_t)(unsigned)((n * H + y) * W + x) * (unsigned)s.C + c0;
      const bool vec = (s.C & 3) == 0;
      load8(s.a + o, cn, vec, v);
      if (s.b) {
        float m[8];
        load8(s.b + o, cn, vec, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= m[i];
      }
      return;
    }
    case CGS_SRC_CATUP: {
      const int C0 = s.C0, C1 = s.C - C0;
      const size_t pa = (size_t)(unsigned)((n * H + y) * W + x) * (unsigned)C0;
      const size_t pb = (size_t)(unsigned)((n * (H >> s.shift) + (y >> s.shift)) * (W >> s.shift) + (x >> s.shift)) * (unsigned)C1;
      if (c0 + cn <= C0) {
        load8(s.a + pa + c0, cn, (C0 & 3) == 0, v);
      } else if (c0 >= C0) {
        load8(s.b + pb + (c0 - C0), cn, ((C1 | (c0 - C0)) & 3) == 0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = c0 + i;
          v[i] = i < cn ? (c < C0 ? __ldg(s.a + pa + c) : __ldg(s.b + pb + (c - C0))) : 0.f;
        }
      }
      return;
    }
    case CGS_SRC_POOLBWD: {
      const size_t o = (size_t)(unsigned)((n * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * (unsigned)s.C + c0;
      const int pos = ((y & 1) << 1) | (x & 1);
      const bool vec = (s.C & 7) == 0;
      float e[8];
      load8(s.a + o, cn, vec, v);
      load8(s.b + o, cn, vec, e);
      unsigned long long ib = 0;
      if (vec && cn == 8) {
        ib = __ldg(reinterpret_cast<const unsigned long long*>(s.idx + o));
      } else {
        for (int i = 0; i < cn; ++i) ib |= (unsigned long long)s.idx[o + i] << (8 * i);
      }
      // ReLU'(pre-act) == (pooled output > 0) at the arg-max position; elsewhere no gradient.
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if ((int)((ib >> (8 * i)) & 0xff) != pos || !(e[i] > 0.f)) v[i] = 0.f;
      return;
    }
    case CGS_SRC_SIGGRAD: {
      const size_t o = (size_t)(unsigned)((n * H + y) * W + x) * (unsigned)s.C + c0;
      const bool vec = (s.C & 3) == 0;
      float z[8];
      load8(s.a + o, cn, vec, v);
      load8(s.b + o, cn, vec, z);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * z[i] * (1.f - z[i]);
      return;
    }
    case CGS_SRC_LEAKYGRAD: {
      const size_t o = (size_t)(unsigned)((n * H + y) * W + x) * (unsigned)s.C + c0;
      const bool vec = (s.C & 3) == 0;
      float z[8];
      load8(s.a + o, cn, vec, v);
      load8(s.b + o, cn, vec, z);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = z[i] > 0.f ? v[i] : v[i] * kLeakySlope;
      return;
    }
    case CGS_SRC_U8ROLL: {
      const int roll = s.b ? *reinterpret_cast<const int*>(s.b) : s.shift;   // |roll| < W
      int sx = x + roll;
      if (sx >= W) sx -= W;
      if (sx < 0) sx += W;
      const uint8_t* q = reinterpret_cast<const uint8_t*>(s.a) + (((size_t)n * H + y) * W + sx) * s.C + c0;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = i < cn ? (float)__ldg(q + i) / 255.0f : 0.f;
      return;
    }
  }
}

// ---- programmatic dependent launch (the wide path's 27-kernel step): a kernel launched through launch_pdl() may become resident
// while the kernel before it in the stream is still draining; pdl_wait() returns once that kernel has completed and its writes
// are visible, so everything before it must touch on-chip state only.  pdl_trigger() lets the NEXT kernel start its own prologue.
// Both are no-ops for plain launches.  CGS_PDL=0 turns the launch attribute off.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

}  // namespace cgs
