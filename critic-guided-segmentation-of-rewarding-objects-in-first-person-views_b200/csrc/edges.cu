// The data formats either side of the hot path, as HBM-bound byte kernels:
//   * cgs_gather_frames      : the contrastive batches `Xpos[Hidx]`, `Xneg[Lidx]`, `Xneg[Cidx]` (reference main.py:345-353) as an
//                              indexed gather over a device-resident uint8 dataset, 16 bytes per thread per access;
//   * cgs_mask_images        : the `-process` outputs (main.py:1212-1223): raw-mask = (M * 255).astype(uint8) and
//                              thresholded-mask = hardM * 255, each replicated to 3 channels, or the `-concatenated` strip
//                              [frame | raw | thresholded] — PNG-ready uint8 rows straight from the mask;
//   * cgs_saliency_normalize : the saliency baseline's normalisation (main.py:974-993 / 1176-1196): per-frame k-th smallest
//                              value (k = int(4096 * thresh)) by a 4-pass radix select in shared memory (or a global norm),
//                              salM = min(sal / norm * pred, 1), salhardM = salM > thresh.
#include "common.cuh"

namespace cgs {

__global__ void gather_frames_kernel(const uint4* __restrict__ data, const int32_t* __restrict__ idx, int n, int64_t nframes,
                                     uint4* __restrict__ out) {
  // one frame = 12288 bytes = 768 uint4; a warp copies 512 contiguous bytes per instruction
  const int64_t total = (int64_t)n * 768;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(e / 768), c = (int)(e - (int64_t)f * 768);
    int64_t s = idx[f];
    s = s < 0 ? 0 : (s >= nframes ? nframes - 1 : s);
    out[e] = __ldg(data + s * 768 + c);
  }
}

__global__ void mask_images_kernel(const float* __restrict__ mask, const uint8_t* __restrict__ hard, const uint8_t* __restrict__ frames,
                                   const uint8_t* __restrict__ lut, int64_t npix, int concatenated, uint8_t* __restrict__ raw,
                                   uint8_t* __restrict__ thr) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    // (M * 255).astype(np.uint8): fp32 product, truncation toward zero (main.py:1216, 1221); hardM * 255 -> 0 / 255
    const uint8_t r = (uint8_t)__float2int_rz(__fmul_rn(__ldg(mask + i), 255.f));
    const uint8_t h = __ldg(hard + i) ? 255 : 0;
    if (concatenated) {
      // strip row layout [64][3 * 64][3]: frame | raw-mask | thresholded-mask (np.concatenate(..., axis=-2), main.py:1216)
      const int64_t f = i >> 12;
      const int y = (int)(i >> 6) & 63, x = (int)i & 63;
      uint8_t* row = raw + ((f * 64 + y) * 192) * 3;
      const uint8_t* src = frames + i * 3;
      // (X / 255.0 * 255).astype(uint8) in float64 is NOT the identity for every byte: the host passes the exact table
      row[x * 3 + 0] = lut[src[0]]; row[x * 3 + 1] = lut[src[1]]; row[x * 3 + 2] = lut[src[2]];
      row[(64 + x) * 3 + 0] = r; row[(64 + x) * 3 + 1] = r; row[(64 + x) * 3 + 2] = r;
      row[(128 + x) * 3 + 0] = h; row[(128 + x) * 3 + 1] = h; row[(128 + x) * 3 + 2] = h;
    } else {
      raw[i * 3 + 0] = r; raw[i * 3 + 1] = r; raw[i * 3 + 2] = r;
      thr[i * 3 + 0] = h; thr[i * 3 + 1] = h; thr[i * 3 + 2] = h;
    }
  }
}

// One CTA (256 threads) per frame: the k-th smallest (0-based, ascending) of the frame's 4096 non-negative floats by an
// MSB-first radix select on the bit patterns (for non-negative floats, integer order == float order), then normalise.
__global__ void __launch_bounds__(256) saliency_normalize_kernel(const float* __restrict__ sal, const float* __restrict__ pred,
                                                                 const float* __restrict__ global_norm, int k, float thresh,
                                                                 float* __restrict__ out, uint8_t* __restrict__ hard,
                                                                 float* __restrict__ norm_out) {
  __shared__ unsigned hist[256];
  __shared__ unsigned sel_prefix, sel_k;
  const int f = blockIdx.x, tid = threadIdx.x;
  const float* s = sal + (size_t)f * 4096;
  unsigned v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__ldg(s + tid + 256 * i));
  float norm;
  if (global_norm) {
    norm = *global_norm;
  } else {
    if (tid == 0) { sel_prefix = 0u; sel_k = (unsigned)k; }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      hist[tid] = 0u;
      __syncthreads();
      const unsigned prefix = sel_prefix, himask = pass ? (0xFFFFFFFFu << (shift + 8)) : 0u;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if ((v[i] & himask) == prefix) atomicAdd(&hist[(v[i] >> shift) & 255u], 1u);
      __syncthreads();
      if (tid == 0) {
        unsigned kk = sel_k, b = 0;
        for (; b < 256; ++b) {
          if (kk < hist[b]) break;
          kk -= hist[b];
        }
        sel_prefix = prefix | (b << shift);
        sel_k = kk;
      }
      __syncthreads();
    }
    norm = __uint_as_float(sel_prefix);
  }
  if (norm_out && tid == 0) norm_out[f] = norm;
  const float pr = __ldg(pred + f);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    // salM / (norm + float_info.min) * pred, clipped at 1 (fp32, as numpy computes it: the tiny addend vanishes in fp32)
    float q = __fmul_rn(__fdiv_rn(__uint_as_float(v[i]), norm), pr);
    if (q >= 1.f) q = 1.f;
    out[(size_t)f * 4096 + tid + 256 * i] = q;
    hard[(size_t)f * 4096 + tid + 256 * i] = q > thresh;
  }
}

int grid_for_edges(int64_t n, int block) {
  const int64_t g = (n + block - 1) / block;
  const int cap = device_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_gather_frames(const uint8_t* dataset, int64_t nframes, const int32_t* idx, int32_t n, uint8_t* out, void* stream) {
  CGS_REQUIRE(dataset && idx && out && n > 0 && nframes > 0, "gather_frames: bad args");
  CGS_REQUIRE((((uintptr_t)dataset | (uintptr_t)out) & 15) == 0, "gather_frames: dataset and out must be 16-byte aligned");
  gather_frames_kernel<<<grid_for_edges((int64_t)n * 768, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(dataset), idx, n, nframes, reinterpret_cast<uint4*>(out));
  return check_launch("gather_frames");
}

extern "C" int cgs_mask_images(const float* mask, const uint8_t* hard, int32_t B, const uint8_t* frames, const uint8_t* lut,
                               int32_t concatenated, uint8_t* raw, uint8_t* thresholded, void* stream) {
  CGS_REQUIRE(mask && hard && raw && B > 0, "mask_images: bad args");
  CGS_REQUIRE(concatenated ? (frames && lut) : (thresholded != nullptr), "mask_images: concatenated needs frames + lut, else two outputs");
  mask_images_kernel<<<grid_for_edges((int64_t)B * 4096, 256), 256, 0, (cudaStream_t)stream>>>(mask, hard, frames, lut, (int64_t)B * 4096,
                                                                                            concatenated, raw, thresholded);
  return check_launch("mask_images");
}

extern "C" int cgs_saliency_normalize(const float* sal, const float* pred, int32_t B, int32_t k, float thresh, const float* global_norm,
                                      float* out, uint8_t* hard, float* norm_out, void* stream) {
  CGS_REQUIRE(sal && pred && out && hard && B > 0, "saliency_normalize: bad args");
  // k = int(salM.shape[-1] * salM.shape[-2] * thresh) (main.py:982), computed by the caller in double precision
  CGS_REQUIRE(global_norm || (k >= 0 && k < 4096), "saliency_normalize: k out of range");
  saliency_normalize_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(sal, pred, global_norm, k, thresh, out, hard, norm_out);
  return check_launch("saliency_normalize");
}
