// Fused NewCritic tail: features[9..15] + crit (reference nets.py:179-195) in ONE kernel each way.
//
//   forward : [Dropout] Conv2d(8c,16c,3,1,1) ReLU MaxPool2d(2) (-> embeds[3]) [Dropout] Conv2d(16c,32c,4) ReLU
//             (-> embeds[4]) Flatten Linear ReLU [Dropout] Linear Sigmoid
//   backward: all of the above, including the 3x3 conv's wgrad + dgrad through ReLU/pool/dropout.
//
// These layers are < 5 % of the critic's MACs but were 6 launches of latency-bound kernels (8x8 and 4x4 maps
// give each CTA almost nothing to do).  Here a CTA owns FPC whole frames; every activation of the tail lives in
// shared memory; only e2 (8x8x8c) is read and embeds[3], embeds[4], pred (+ what backward needs) are written.
// Weight gradients are reduced over the CTA's frames and pushed with one RED per weight per CTA.
#include <type_traits>
#include "tail_body.cuh"

namespace cgs {

template <int FPC>
__global__ void __launch_bounds__(HT) tail_fwd_kernel(const TailDims d, const float* __restrict__ e2, const float* __restrict__ m_e2,
                                                      const float* __restrict__ m_e3, const float* __restrict__ m_v,
                                                      const float* __restrict__ w3, const float* __restrict__ b3,
                                                      const float* __restrict__ w14, const float* __restrict__ b14,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      float* __restrict__ e3, uint8_t* __restrict__ idx3, float* __restrict__ e4,
                                                      float* __restrict__ vout, float* __restrict__ pred) {
  extern __shared__ __align__(16) float smem[];
  tail_fwd_body<FPC>(smem, blockIdx.x * FPC, false, d, e2, m_e2, m_e3, m_v, w3, b3, w14, b14, w1, b1, w2, b2, e3, idx3, e4, vout, pred);
}

template <int FPC>
__global__ void __launch_bounds__(HT) tail_bwd_kernel(const TailDims d, const float* __restrict__ e2, const float* __restrict__ m_e2,
                                                      const float* __restrict__ m_e3, const float* __restrict__ m_v,
                                                      const float* __restrict__ w3, const float* __restrict__ w14,
                                                      const float* __restrict__ w1, const float* __restrict__ w2,
                                                      const float* __restrict__ e3, const uint8_t* __restrict__ idx3,
                                                      const float* __restrict__ e4, const float* __restrict__ v,
                                                      const float* __restrict__ pred, const float* __restrict__ dpred,
                                                      const float* __restrict__ de3_ext, const float* __restrict__ de4,
                                                      const TailGrads G, float* __restrict__ de2) {
  extern __shared__ __align__(16) float smem[];
  tail_bwd_body<FPC>(smem, blockIdx.x * FPC, d, e2, m_e2, m_e3, m_v, w3, w14, w1, w2, e3, idx3, e4, v, pred, dpred, de3_ext, de4, G,
                     de2, nullptr);
}

static size_t tail_bytes(int FPC, int C2, int C3, int NB, bool bwd) { return (size_t)tail_smem(FPC, C2, C3, NB, bwd).total * sizeof(float); }
static int tail_fpc(int B) { return B <= 512 ? 1 : (B <= 4096 ? 2 : 8); }

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_tail_supported(int32_t B, int32_t C2, int32_t C3, int32_t NB) {
  return tail_bytes(tail_fpc(B), C2, C3, NB, true) <= 160 * 1024 ? 1 : 0;
}

extern "C" int cgs_tail_fwd(const float* e2, const float* m_e2, const float* m_e3, const float* m_v, const float* w3,
                            const float* b3, const float* w14, const float* b14, const float* w1, const float* b1,
                            const float* w2, const float* b2, int32_t B, int32_t C2, int32_t C3, int32_t NB, float* e3,
                            uint8_t* idx3, float* e4, float* v, float* pred, void* stream) {
  CGS_REQUIRE(e2 && w3 && b3 && w14 && b14 && w1 && b1 && w2 && b2 && e3 && idx3 && e4 && v && pred, "tail_fwd: null pointer");
  CGS_REQUIRE(B > 0 && C2 > 0 && C3 > 0 && NB > 0, "tail_fwd: bad sizes");
  CGS_REQUIRE(cgs_tail_supported(B, C2, C3, NB), "tail_fwd: shapes exceed the fused kernel's shared-memory budget");
  const TailDims d{B, C2, C3, NB};
  auto go = [&](auto tag) {
    constexpr int FPC = decltype(tag)::value;
    const size_t smem = tail_bytes(FPC, C2, C3, NB, false);
    if (smem > 48 * 1024) cudaFuncSetAttribute(tail_fwd_kernel<FPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tail_fwd_kernel<FPC><<<(B + FPC - 1) / FPC, HT, smem, (cudaStream_t)stream>>>(d, e2, m_e2, m_e3, m_v, w3, b3, w14, b14, w1, b1,
                                                                                 w2, b2, e3, idx3, e4, v, pred);
    return check_launch("tail_fwd");
  };
  const int fpc = tail_fpc(B);
  return fpc == 1 ? go(std::integral_constant<int, 1>{}) : fpc == 2 ? go(std::integral_constant<int, 2>{}) : go(std::integral_constant<int, 8>{});
}

extern "C" int cgs_tail_bwd(const float* e2, const float* m_e2, const float* m_e3, const float* m_v, const float* w3,
                            const float* w14, const float* w1, const float* w2, const float* e3, const uint8_t* idx3,
                            const float* e4, const float* v, const float* pred, const float* dpred, const float* de3_ext,
                            const float* de4, int32_t B, int32_t C2, int32_t C3, int32_t NB, float* dw3, float* db3,
                            float* dw14, float* db14, float* dw1, float* db1, float* dw2, float* db2, float* de2,
                            void* stream) {
  CGS_REQUIRE(e2 && w3 && w14 && w1 && w2 && e3 && idx3 && e4 && v && pred, "tail_bwd: null pointer");
  CGS_REQUIRE(dpred || de3_ext || de4, "tail_bwd: no incoming gradient");
  CGS_REQUIRE(B > 0 && C2 > 0 && C3 > 0 && NB > 0, "tail_bwd: bad sizes");
  if (dw14) CGS_REQUIRE(dw3 && db3 && db14 && dw1 && db1 && dw2 && db2, "tail_bwd: parameter gradients are all-or-none");
  CGS_REQUIRE(cgs_tail_supported(B, C2, C3, NB), "tail_bwd: shapes exceed the fused kernel's shared-memory budget");
  const TailDims d{B, C2, C3, NB};
  const TailGrads G{dw14 ? dw3 : nullptr, db3, dw14, db14, dw1, db1, dw2, db2};
  auto go = [&](auto tag) {
    constexpr int FPC = decltype(tag)::value;
    const size_t smem = tail_bytes(FPC, C2, C3, NB, true);
    if (smem > 48 * 1024) cudaFuncSetAttribute(tail_bwd_kernel<FPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tail_bwd_kernel<FPC><<<(B + FPC - 1) / FPC, HT, smem, (cudaStream_t)stream>>>(d, e2, m_e2, m_e3, m_v, w3, w14, w1, w2, e3, idx3,
                                                                                 e4, v, pred, dpred, de3_ext, de4, G, de2);
    return check_launch("tail_bwd");
  };
  const int fpc = tail_fpc(B);
  return fpc == 1 ? go(std::integral_constant<int, 1>{}) : fpc == 2 ? go(std::integral_constant<int, 2>{}) : go(std::integral_constant<int, 8>{});
}
