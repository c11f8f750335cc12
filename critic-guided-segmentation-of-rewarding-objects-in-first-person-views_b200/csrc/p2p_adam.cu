// Data-parallel gradient exchange fused with the optimizer: a one-shot all-reduce over NVLink peer memory + Adam.
//
// The gradient bucket of this path is 47-103 KB: an NCCL all-reduce of that size is pure latency (~20 us inside a
// 70-90 us step).  Here every rank keeps its summed local gradient in a SYMMETRIC buffer (same allocation mapped into
// every peer, torch.distributed._symmetric_memory); one kernel per rank
//   1. tells every peer "my gradient of step t is complete" (one 32-bit store per peer into the peer's flag pad),
//   2. waits until every peer has said the same (spins on LOCAL memory, bounded),
//   3. sums element i over all ranks' buffers in rank order (peer loads over NVLink/NVSwitch; the same order on every rank
//      -> bit-identical parameters everywhere, no broadcast needed),
//   4. applies Adam to its own copy of the parameters.
// Two gradient slots alternate by step parity and the flags only ever grow (flag = step number), so no second barrier and
// no flag reset is needed: a rank can run at most one step ahead of a peer, and then it writes the other slot.
#include "common.cuh"

namespace cgs {

constexpr int P2P_MAX_WORLD = 16;

struct P2PPtrs {
  const float* buf[P2P_MAX_WORLD];   // every rank's gradient buffer [2][npad], as mapped into THIS process
  unsigned* flag[P2P_MAX_WORLD];     // every rank's flag pad [P2P_MAX_WORLD]
};

// stage: sym[slot(t)][i] = g[i] + sum_k partials[k][i - off]  (fixed order); g is cleared.  Same block shape as
// partials_kernel (elementwise.cu); np == 0 just moves the bucket.
__global__ void __launch_bounds__(256) p2p_stage_kernel(float* __restrict__ g, int64_t n, int64_t npad, float* __restrict__ sym,
                                                        const float* __restrict__ part, int np, int64_t stride, int64_t off,
                                                        int64_t len, const int* __restrict__ step_state) {
  __shared__ float red[8][33];
  const int ex = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + ex;
  const int64_t j = i - off;
  float s0 = 0.f;
  if (np > 0 && i < n && j >= 0 && j < len) {
    const float* q = part + j;
    for (int k0 = ky; k0 < np; k0 += 128) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (k0 + 8 * u < np) ? __ldg(q + (int64_t)(k0 + 8 * u) * stride) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) s0 += v[u];
    }
  }
  red[ky][ex] = s0;
  __syncthreads();
  if (ky == 0 && i < n) {
    float gv = g[i];
#pragma unroll
    for (int r = 0; r < 8; ++r) gv += red[r][ex];
    const int ep = step_state[2] + 1;             // exchange epoch (monotonic; see p2p_allreduce_adam_kernel)
    sym[(int64_t)(ep & 1) * npad + i] = gv;
    g[i] = 0.f;
  }
}

__global__ void __launch_bounds__(256) p2p_allreduce_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                                 int64_t n, int64_t npad, const P2PPtrs pp, int rank, int world,
                                                                 double lr, double beta1, double beta2, double eps_d,
                                                                 int* __restrict__ step_state, float gscale, int* __restrict__ err) {
  const int t = step_state[0] + 1;
  // the flags and the slot parity follow the exchange epoch, which only ever grows; the Adam step count t may be rewound
  // (CUDA-graph capture restores it after its warm-up steps) and must not be what peers wait for
  const int ep = step_state[2] + 1;
  __shared__ int s_failed;
  if (threadIdx.x == 0) s_failed = 0;
  __syncthreads();
  // 1. announce (block 0 only; the staging kernel before us on this stream has completed, its stores are in our L2)
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(pp.flag[threadIdx.x] + rank), "r"((unsigned)ep) : "memory");
  }
  // 2. wait for every peer's announcement of step t (bounded: a diverged peer must not hang the GPU)
  if (threadIdx.x < world) {
    const unsigned* f = pp.flag[rank] + threadIdx.x;
    unsigned seen = 0;
    for (long long spin = 0; spin < (1ll << 26); ++spin) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(f) : "memory");
      if ((int)seen >= ep) break;
      __nanosleep(64);
    }
    if ((int)seen < ep) { atomicExch(err, 1); s_failed = 1; }
  }
  __shared__ float s_bc[2];
  if (threadIdx.x == 32) {                         // one double-precision pow pair per block, while the others poll
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    s_bc[0] = (float)(lr / bc1);
    s_bc[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  // 3. + 4.
  const float step_size = s_bc[0];
  const float bc2_sqrt = s_bc[1];
  const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2), eps = (float)eps_d;
  const int64_t slot = (int64_t)(ep & 1) * npad;
  // a peer that never announced leaves an incomplete sum: keep parameters and moments as they are (flag is set)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n && !s_failed; i += (int64_t)gridDim.x * blockDim.x) {
    float gv = 0.f;
    for (int r = 0; r < world; ++r) {
      float x;
      asm volatile("ld.relaxed.sys.global.f32 %0, [%1];\n" : "=f"(x) : "l"(pp.buf[r] + slot + i) : "memory");
      gv += x;
    }
    gv *= gscale;
    const float mv = m[i] + omb1 * (gv - m[i]);
    const float vv = v[i] * b2 + omb2 * gv * gv;
    m[i] = mv;
    v[i] = vv;
    p[i] -= step_size * (mv / (sqrtf(vv) / bc2_sqrt + eps));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&step_state[1], 1) == (int)gridDim.x - 1) {
      step_state[1] = 0;
      step_state[0] = t;
      step_state[2] = ep;
    }
  }
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_p2p_stage(float* g, int64_t n, int64_t npad, float* sym_local, const float* partials, int32_t n_partials,
                             int64_t stride, int64_t offset, int64_t len, const int32_t* step_state, void* stream) {
  CGS_REQUIRE(g && sym_local && step_state && n > 0 && npad >= n, "p2p_stage: bad args");
  CGS_REQUIRE(n_partials == 0 || (partials && offset >= 0 && offset + len <= n), "p2p_stage: bad partials");
  p2p_stage_kernel<<<(unsigned)((n + 31) / 32), 256, 0, (cudaStream_t)stream>>>(g, n, npad, sym_local, partials, n_partials, stride,
                                                                             offset, len, step_state);
  return check_launch("p2p_stage");
}

extern "C" int cgs_p2p_allreduce_adam(float* p, float* m, float* v, int64_t n, int64_t npad, const uint64_t* peer_bufs,
                                      const uint64_t* peer_flags, int32_t rank, int32_t world, double lr, double beta1,
                                      double beta2, double eps, int32_t* step_state, float grad_scale, int32_t* err_flag,
                                      void* stream) {
  CGS_REQUIRE(p && m && v && peer_bufs && peer_flags && step_state && err_flag && n > 0 && npad >= n, "p2p_allreduce_adam: bad args");
  CGS_REQUIRE(world >= 2 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "p2p_allreduce_adam: world %d rank %d", world, rank);
  P2PPtrs pp;
  for (int r = 0; r < P2P_MAX_WORLD; ++r) {
    pp.buf[r] = r < world ? reinterpret_cast<const float*>(peer_bufs[r]) : nullptr;
    pp.flag[r] = r < world ? reinterpret_cast<unsigned*>(peer_flags[r]) : nullptr;
  }
  // every block spins until the peers have announced: keep the grid small enough to be co-resident on any device state
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > 128) grid = 128;
  p2p_allreduce_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, m, v, n, npad, pp, rank, world, lr, beta1, beta2, eps,
                                                                   step_state, grad_scale, err_flag);
  return check_launch("p2p_allreduce_adam");
}
