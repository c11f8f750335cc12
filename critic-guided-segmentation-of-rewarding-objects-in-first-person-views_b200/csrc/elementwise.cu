// HBM-bound stages of the hot path: occlusion blend (main.py:395,406) forward/backward,
// prediction losses (main.py:193-195,400,411), mask regulariser (main.py:415-429),
// uint8->float frame conversion with the shift_batch roll (main.py:189,584-591), binary
// threshold (main.py:964,1164) and the flat Adam update (main.py:178,331-334).
// All use grid-stride loops over 128-bit accesses where alignment allows, warp-shuffle +
// one-atomic-per-CTA reductions, and grids sized to a multiple of the SM count.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

namespace cgs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return CGS_ECUDA;
  }
  return CGS_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CGS_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int device_sms() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  int sms = cache[slot];
  if (sms == 0) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cache[slot] = sms;     // benign race: every thread computes the same value
  }
  return sms;
}

static int grid_for(int64_t work_items, int threads) {
  const int sms = device_sms();
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float s_part[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x + 31) / 32 ? s_part[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// ---------------------------------------------------------------- occlusion blend
// One thread per pixel (C contiguous floats); C == 3 on the path.  Loads of consecutive
// pixels are contiguous across the warp (3 x 128 B per 32 pixels per tensor).
__global__ void occlude_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ z,
                                   int64_t npix, int C, float* __restrict__ out) {
  const int64_t total = npix * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float zz = __ldg(z + i / C);
    const float av = __ldg(a + i), bv = __ldg(b + i);
    out[i] = av * (1.f - zz) + zz * bv;   // same association as the reference expression
  }
}

__global__ void occlude_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ z,
                                   const float* __restrict__ g, int64_t npix, int C, float* __restrict__ dz,
                                   float* __restrict__ da, float* __restrict__ db) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const float zz = __ldg(z + p);
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
      const int64_t i = p * C + c;
      const float gv = __ldg(g + i);
      acc = fmaf(__ldg(b + i) - __ldg(a + i), gv, acc);
      if (da) da[i] = gv * (1.f - zz);
      if (db) db[i] = gv * zz;
    }
    if (dz) dz[p] = acc;
  }
}

// ---------------------------------------------------------------- prediction loss
__global__ void pred_loss_kernel(const float* __restrict__ p, const float* __restrict__ t, int n, int bce, float gscale,
                                 float* __restrict__ loss, float* __restrict__ grad) {
  float acc = 0.f;
  const float inv = 1.f / (float)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float pv = __ldg(p + i), tv = __ldg(t + i);
    if (bce) {
      // F.binary_cross_entropy clamps the logs at -100
      const float lp = fmaxf(logf(pv), -100.f), l1p = fmaxf(logf(1.f - pv), -100.f);
      acc -= tv * lp + (1.f - tv) * l1p;
      if (grad) grad[i] = gscale * inv * (pv - tv) / fmaxf(pv * (1.f - pv), 1e-12f);
    } else {
      const float d = pv - tv;
      acc = fmaf(d, d, acc);
      if (grad) grad[i] = gscale * inv * 2.f * d;
    }
  }
  const float r = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, r * inv);
}

// ---------------------------------------------------------------- mask regulariser
__global__ void mask_reg_kernel(const float* __restrict__ z, const float* __restrict__ vpred, int64_t n, int per_frame,
                                float l1, float l2, float gscale, float* __restrict__ loss, float* __restrict__ grad) {
  float acc = 0.f;
  const float inv = 1.f / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float vf = vpred ? 1.f - __ldg(vpred + i / per_frame) : 1.f;
    const float u = vf * __ldg(z + i);
    acc += l1 * fabsf(u) + l2 * u * u;
    if (grad) {
      const float sg = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
      grad[i] = gscale * inv * vf * (l1 * sg + 2.f * l2 * u);
    }
  }
  const float r = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, r * inv);
}

// ---------------------------------------------------------------- frames uint8 -> float
// 4 consecutive output floats (one 128-bit store) per thread; the roll only permutes bytes inside a row of
// W*C bytes, so the source is 4 byte loads from the same (L1-resident) row.  Division-free (FastDiv by W*C/4).
__global__ void frames_to_float_kernel(const uint8_t* __restrict__ in, int64_t nvec, int rowc, FastDiv dq, int rollc,
                                       const int* __restrict__ roll_dev, int C, float* __restrict__ out) {
  if (roll_dev) {
    int r = *roll_dev * C % rowc;
    rollc = r < 0 ? r + rowc : r;
  }
  const int q = rowc >> 2;   // float4 per row
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = (i < 0x7fffffff / q) ? (int64_t)fdiv((int)i, dq) : i / q;
    const int j = (int)(i - row * q) * 4;
    const uint8_t* src = in + row * rowc;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int o = j + k + rollc;
      if (o >= rowc) o -= rowc;
      v[k] = (float)__ldg(src + o) / 255.0f;
    }
    *reinterpret_cast<float4*>(out + row * rowc + j) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void threshold_kernel(const float* __restrict__ z, int64_t n, float thresh, int strict, uint8_t* __restrict__ hard) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(z + i);
    hard[i] = strict ? (v > thresh) : (v >= thresh);
  }
}

// ---------------------------------------------------------------- dropout masks
// All Dropout masks of one NewCritic forward (nets.py:179,183,192) in ONE launch: Philox4x32-10 keyed by
// (seed, call counter), 4 Bernoulli draws per counter value; value = keep ? 1/(1-p) : 0.  The call counter lives in
// device memory and is advanced by the last CTA to finish (ticket), so a captured CUDA graph draws fresh masks on
// every replay without any host involvement.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void dropout_masks_kernel(float* __restrict__ out, int64_t n, float p, unsigned long long seed,
                                     unsigned long long* __restrict__ state) {
  const unsigned long long call = state[0];
  const float keep = 1.f / (1.f - p);
  const int64_t nvec = (n + 3) >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ((float)(c[k] >> 8) * (1.0f / 16777216.0f) >= p) ? keep : 0.f;
    if (4 * i + 3 < n) {
      *reinterpret_cast<float4*>(out + 4 * i) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int k = 0; k < 4 && 4 * i + k < n; ++k) out[4 * i + k] = v[k];
    }
  }
  // last CTA to finish advances the call counter: every CTA has read state[0] by then
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(&state[1], 1ull);
    if (t == gridDim.x - 1) {
      state[1] = 0;
      state[0] = call + 1;
    }
  }
}

// ---------------------------------------------------------------- Adam
// step_state = {step counter t (already-applied steps), ticket}: the kernel applies step t+1 and the last CTA to finish
// publishes t+1, so no separate counter-increment launch is needed.  With clear_grad the consumed gradient is zeroed in
// the same pass (next step's wgrad kernels accumulate from zero: no memset launch).
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, double lr, double beta1, double beta2, double eps_d,
                            int* __restrict__ step_state, float gscale, int clear_grad) {
  const int t = step_state[0] + 1;
  // bias corrections in double, as torch does on the host (python floats); tensor-side scalars in fp32.
  // One thread per block does the double-precision pow, the rest read the two results from shared memory.
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    s_bc[0] = (float)(lr / bc1);
    s_bc[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_bc[0];
  const float bc2_sqrt = s_bc[1];
  const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2), eps = (float)eps_d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gv = g[i] * gscale;
    const float mv = m[i] + omb1 * (gv - m[i]);                 // lerp form used by torch
    const float vv = v[i] * b2 + omb2 * gv * gv;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] -= step_size * (mv / denom);
    if (clear_grad) g[i] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&step_state[1], 1) == (int)gridDim.x - 1) {
      step_state[1] = 0;
      step_state[0] = t;
    }
  }
}

// ---------------------------------------------------------------- -eval: hard mask vs ground truth -> IoU counts
// hardM = M > eval_thresh (reference main.py:964; >= for the -process convention), get_iou's np.sum(A & B), np.sum(A | B)
// (main.py:1265-1270) as one pass: compare, ballot, popcount, one 64-bit atomic pair per CTA.  Exact integer arithmetic.
__global__ void iou_counts_kernel(const float* __restrict__ z, const uint8_t* __restrict__ gt, int64_t n, float thresh, int strict,
                                  unsigned long long* __restrict__ counts) {
  unsigned inter = 0, uni = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(z + i);
    const bool h = strict ? (v > thresh) : (v >= thresh), g = __ldg(gt + i) != 0;
    inter += h && g;
    uni += h || g;
  }
  __shared__ unsigned s_i[8], s_u[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { inter += __shfl_xor_sync(0xffffffffu, inter, o); uni += __shfl_xor_sync(0xffffffffu, uni, o); }
  if ((threadIdx.x & 31) == 0) { s_i[threadIdx.x >> 5] = inter; s_u[threadIdx.x >> 5] = uni; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned a = 0, b = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_i[w]; b += s_u[w]; }
    atomicAdd(counts, (unsigned long long)a);
    atomicAdd(counts + 1, (unsigned long long)b);
  }
}

// Gradient = bucket + the sum of per-CTA partial gradient vectors (written by the whole-step critic kernel instead of
// same-address REDs), reduced in a FIXED order (bit-reproducible), then either written back to the bucket (data-parallel:
// the all-reduce comes next) or consumed by Adam in the same pass.  Block = 32 elements x 8 slices of the partial list:
// every warp reads 128 contiguous bytes per partial, eight partials in flight per thread.
template <bool ADAM>
__global__ void __launch_bounds__(256) partials_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, int64_t n, const float* __restrict__ part, int np,
                                                       int64_t stride, int64_t off, int64_t len, double lr, double beta1,
                                                       double beta2, double eps_d, int* __restrict__ step_state, float gscale) {
  __shared__ float red[8][33];
  const int ex = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + ex;
  const int64_t j = i - off;
  float s0 = 0.f;
  if (i < n && j >= 0 && j < len) {
    const float* q = part + j;
    for (int k0 = ky; k0 < np; k0 += 128) {        // 16 loads in flight per thread: one L2 round trip for <= 128 partials
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (k0 + 8 * u < np) ? __ldg(q + (int64_t)(k0 + 8 * u) * stride) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) s0 += v[u];
    }
  }
  red[ky][ex] = s0;
  __shared__ float s_bc[2];
  if (ADAM && threadIdx.x == 0) {
    const int t = step_state[0] + 1;
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    s_bc[0] = (float)(lr / bc1);
    s_bc[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  if (ky == 0 && i < n) {
    float gv = g[i];
#pragma unroll
    for (int r = 0; r < 8; ++r) gv += red[r][ex];
    if (!ADAM) {
      g[i] = gv;
    } else {
      const float step_size = s_bc[0];
      const float bc2_sqrt = s_bc[1];
      const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2), eps = (float)eps_d;
      gv *= gscale;
      const float mv = m[i] + omb1 * (gv - m[i]);
      const float vv = v[i] * b2 + omb2 * gv * gv;
      m[i] = mv;
      v[i] = vv;
      p[i] -= step_size * (mv / (sqrtf(vv) / bc2_sqrt + eps));
      g[i] = 0.f;
    }
  }
  if (ADAM) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int t = step_state[0] + 1;
      __threadfence();
      if (atomicAdd(&step_state[1], 1) == (int)gridDim.x - 1) {
        step_state[1] = 0;
        step_state[0] = t;
      }
    }
  }
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_iou_counts(const float* z, const uint8_t* gt, int64_t n, float thresh, int32_t strict, uint64_t* counts,
                              void* stream) {
  CGS_REQUIRE(z && gt && counts && n > 0, "iou_counts: bad args");
  iou_counts_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(z, gt, n, thresh, strict, (unsigned long long*)counts);
  return check_launch("iou_counts");
}

extern "C" int cgs_reduce_partials(float* g, int64_t n, const float* partials, int32_t n_partials, int64_t stride,
                                   int64_t offset, int64_t len, void* stream) {
  CGS_REQUIRE(g && partials && n > 0 && n_partials > 0 && offset >= 0 && offset + len <= n, "reduce_partials: bad args");
  partials_kernel<false><<<(unsigned)((n + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
      nullptr, g, nullptr, nullptr, n, partials, n_partials, stride, offset, len, 0, 0, 0, 0, nullptr, 1.f);
  return check_launch("reduce_partials");
}

extern "C" int cgs_adam_step_partials(float* p, float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                                      double eps, int32_t* step_state, float grad_scale, const float* partials,
                                      int32_t n_partials, int64_t stride, int64_t offset, int64_t len, void* stream) {
  CGS_REQUIRE(p && g && m && v && step_state && n > 0 && partials && n_partials > 0 && offset >= 0 && offset + len <= n,
              "adam_step_partials: bad args");
  partials_kernel<true><<<(unsigned)((n + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, partials, n_partials, stride, offset, len, lr, beta1, beta2, eps, step_state, grad_scale);
  return check_launch("adam_step_partials");
}

extern "C" int cgs_occlude_fwd(const float* a, const float* b, const float* z, int64_t npix, int32_t C, float* out,
                               void* stream) {
  CGS_REQUIRE(a && b && z && out && npix > 0 && C > 0, "occlude_fwd: bad args");
  occlude_fwd_kernel<<<grid_for(npix * C, 256), 256, 0, (cudaStream_t)stream>>>(a, b, z, npix, C, out);
  return check_launch("occlude_fwd");
}

extern "C" int cgs_occlude_bwd(const float* a, const float* b, const float* z, const float* g, int64_t npix, int32_t C,
                               float* dz, float* da, float* db_, void* stream) {
  CGS_REQUIRE(a && b && z && g && npix > 0 && C > 0 && (dz || da || db_), "occlude_bwd: bad args");
  occlude_bwd_kernel<<<grid_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(a, b, z, g, npix, C, dz, da, db_);
  return check_launch("occlude_bwd");
}

extern "C" int cgs_pred_loss(const float* p, const float* t, int32_t n, int32_t bce, float gscale, float* loss,
                             float* grad, void* stream) {
  CGS_REQUIRE(p && t && loss && n > 0, "pred_loss: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) return check_launch("pred_loss.memset");
  pred_loss_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, t, n, bce, gscale, loss, grad);
  return check_launch("pred_loss");
}

extern "C" int cgs_mask_reg(const float* z, const float* vpred, int64_t n, int32_t per_frame, float l1, float l2,
                            float gscale, float* loss, float* grad, void* stream) {
  CGS_REQUIRE(z && loss && n > 0 && per_frame > 0, "mask_reg: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) return check_launch("mask_reg.memset");
  mask_reg_kernel<<<grid_for(n, 256), 256, 0, st>>>(z, vpred, n, per_frame, l1, l2, gscale, loss, grad);
  return check_launch("mask_reg");
}

extern "C" int cgs_frames_to_float(const uint8_t* in, int32_t B, int32_t H, int32_t W, int32_t C, int32_t roll,
                                   const int32_t* roll_dev, float* out, void* stream) {
  CGS_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0, "frames_to_float: bad args");
  const int rowc = W * C;
  CGS_REQUIRE((rowc & 3) == 0 && rowc >= 8, "frames_to_float: W*C must be a multiple of 4 (got %d)", rowc);
  int rollc = (int)(((int64_t)roll * C) % rowc);
  if (rollc < 0) rollc += rowc;
  const int64_t nvec = (int64_t)B * H * (rowc / 4);
  frames_to_float_kernel<<<grid_for(nvec, 256), 256, 0, (cudaStream_t)stream>>>(in, nvec, rowc, make_fastdiv(rowc / 4), rollc,
                                                                             roll_dev, C, out);
  return check_launch("frames_to_float");
}

extern "C" int cgs_threshold(const float* z, int64_t n, float thresh, int32_t strict, uint8_t* hard, void* stream) {
  CGS_REQUIRE(z && hard && n > 0, "threshold: bad args");
  threshold_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(z, n, thresh, strict, hard);
  return check_launch("threshold");
}

extern "C" int cgs_adam_step(float* p, float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                             double eps, int32_t* step_state, float grad_scale, int32_t clear_grad, void* stream) {
  CGS_REQUIRE(p && g && m && v && step_state && n > 0, "adam_step: bad args");
  adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_state, grad_scale,
                                                              clear_grad);
  return check_launch("adam_step");
}

extern "C" int cgs_dropout_masks(float* out, int64_t n, float p, uint64_t seed, uint64_t* state, void* stream) {
  CGS_REQUIRE(out && state && n > 0 && p >= 0.f && p < 1.f, "dropout_masks: bad args");
  CGS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "dropout_masks: out must be 16-byte aligned");
  dropout_masks_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, p, (unsigned long long)seed,
                                                                                    (unsigned long long*)state);
  return check_launch("dropout_masks");
}

extern "C" const char* cgs_last_error(void) { return g_err; }
extern "C" int cgs_version(void) { return 100; }
