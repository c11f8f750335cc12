// Shared device/host helpers for the cgs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "cgs_b200.h"

namespace cgs {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

}  // namespace cgs
