// What the two whole-step critic kernels (critic_fused.cu: TF32, hg_critic.cu: bf16) share: the parameter block, the layout
// of the per-CTA gradient accumulators / partial vectors, and the tail of a step - gradient hand-over, grid barrier, slice
// reduction, [all-reduce over NVLink peer memory], Adam.
#pragma once
#include "fused_common.cuh"

namespace cgs {
namespace cf {

constexpr int NT_TAIL = 512;
// gradient accumulators of a CTA (float offsets): everything but features.14.weight, which lives in registers (16 per thread)
constexpr int aW0 = 0, aB0 = 216, aW1 = 224, aB1 = 800, aW2 = 808, aB2 = 1384, aW3 = 1392, aB3 = 2544, aB4 = 2560,
              aWl1 = 2592, aBl1 = 3616, aWl2 = 3648, aBl2 = 3680, szAcc = 3712;
constexpr int NGRAD = 11873;                       // critic parameters in state_dict order (= flat gradient layout)
constexpr int PSTRIDE = 11904;                     // per-CTA stride of the partial-gradient buffer (16-byte multiple)

struct Params {
  const uint8_t* frames;
  // input-gradient mode (template XG): fp32 NHWC frames in, d loss / d frame out, frozen parameters (no weight gradients)
  const float* xin;
  float* dx;
  const float* target;
  const float *m2, *m3, *mv;
  const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4, *wl1, *bl1, *wl2, *bl2;
  float* gseg[14];      // gradient tensors in sAcc order (w0 b0 w1 b1 w2 b2 w3 b3 b4 wl1 bl1 wl2 bl2) + [13] = w4
  unsigned long long seed;          // rng_state != NULL: dropout masks are drawn in the kernel (Philox4x32-10, the stream of
  unsigned long long* rng_state;    // cgs_dropout_masks: counter = (float4 index in the [m2|m3|mv] buffer, call), key = seed)
  float p_drop, keep;
  // adam_p != NULL (single GPU, the bucket is exactly the critic): Adam runs in this kernel behind a grid barrier
  float *adam_p, *adam_g, *adam_m, *adam_v;
  double lr, beta1, beta2, eps;
  int* step_state;
  unsigned* bar;        // {arrivals, generation, timed-out flag}, zeroed once by the host
  // world > 1: the gradient is all-reduced inside this kernel over NVLink peer memory, low-latency style: every rank owns
  // a symmetric receive buffer recv[2 slots][world][npad] of {value, step} pairs.  CTA b PUSHES each element of its slice
  // of this rank's summed gradient, tagged with the step number, into every peer's buffer with one 8-byte store (value and
  // tag travel together: no fence, no separate flag), then polls its OWN buffer until all `world` tags of an element
  // equal the step, sums in rank order (bit-identical parameters on every rank) and applies Adam.  One launch per
  // data-parallel step; slots alternate by step parity and tags only grow, so nothing is ever reset.
  int world, rank;
  long long npad;
  unsigned long long* ll_peer[16];  // every rank's receive buffer as mapped into this process
  float* partials;      // != NULL: per-CTA partial gradients [grid][PSTRIDE] in flat order instead of REDs into gseg
  // MODE 3 (the two scored blends of one Hourglass step, main.py:395-411): pass 0 scores replaced = A(1-Z) + Z*B against
  // `target` (negpred), pass 1 scores injected = B(1-Z) + Z*A against `target2` (pred of critic(A)); the blends are formed
  // in shared memory from the raw uint8 frames and the mask, and the input gradient is contracted with (B - A) / (A - B)
  // on the way out: dz = d(replace + inject + regulariser) / dZ.  Neither blend nor its gradient ever exists in HBM.
  const uint8_t* framesB;           // contrast frames (never rolled)
  const float* zmask;               // [B][64][64]
  const float* target2;             // NULL: no inject pass
  const float *m2b, *m3b, *mvb;     // forced dropout masks of pass 1
  float* pred2;
  float* dz;                        // [B][64][64]
  const float* vpred;               // regulariser weight vf = 1 - vpred[n] (non-static norm); NULL: vf = 1
  float l1, l2, reg_scale;          // main.py:415-429; reg_scale = loss_grad / (B * 4096)
  float* pred;
  float* loss;
  const int* roll_dev;
  int B, roll, bce;
  float gscale;         // d(total loss)/d(this rank's mean loss) / B
  float inv_n;          // 1 / B
};

// This CTA's gradient: one coalesced partial vector (summed by the Adam pass), or REDs into the caller's gradient tensors.
__device__ __forceinline__ void grad_handover(const Params& p, const float* sAcc, const float (&accW4)[16], int tid) {
  constexpr int NT = NT_TAIL;
    if (p.partials) {
      // flat (state_dict) order: sAcc[0, 2560) -> [0, 2560); features.14.weight -> [2560, 10752); sAcc[2560, 3681) -> +8192
      float* d = p.partials + (size_t)blockIdx.x * PSTRIDE;
      for (int e = tid; e < 3681; e += NT) d[e < 2560 ? e : e + 8192] = sAcc[e];
      float4* d4 = reinterpret_cast<float4*>(d + 2560 + (tid >> 4) * 256 + (tid & 15) * 16);
      const int rot4 = (tid >> 1) & 3;
  #pragma unroll
      for (int i = 0; i < 4; ++i) d4[(i + rot4) & 3] = make_float4(accW4[4 * i], accW4[4 * i + 1], accW4[4 * i + 2], accW4[4 * i + 3]);
    } else {
      constexpr int segoff[13] = {aW0, aB0, aW1, aB1, aW2, aB2, aW3, aB3, aB4, aWl1, aBl1, aWl2, aBl2};
      constexpr int seglen[13] = {216, 8, 576, 8, 576, 8, 1152, 16, 32, 1024, 32, 32, 1};
  #pragma unroll
      for (int s = 0; s < 13; ++s) {
        float* d = p.gseg[s];
        for (int e = tid; e < seglen[s]; e += NT) atomicAdd(d + e, sAcc[segoff[s] + e]);
      }
      float* d4 = p.gseg[13] + (tid >> 4) * 256 + (tid & 15) * 16;
      const int rot4 = (tid >> 1) & 3;
  #pragma unroll
      for (int i = 0; i < 16; ++i) atomicAdd(d4 + (((i >> 2) + rot4) & 3) * 4 + (i & 3), accW4[i]);
    }
  }

// Grid barrier (all CTAs co-resident), then every CTA sums its slice of the parameter vector over all partial vectors in a
// fixed order, exchanges it with the peers (world > 1) and applies Adam.  scratch: >= 523 floats of shared memory.
template <class Mark>
__device__ __forceinline__ void adam_tail(const Params& p, float* scratch, unsigned bar_gen, int tid, int warp, int lane, Mark&& mark) {
  constexpr int NT = NT_TAIL;
  (void)NT;
    // ---- grid barrier (all CTAs are co-resident: one per SM, grid <= SMs), then every CTA sums its slice of the
    //      parameter vector over all partial vectors (fixed order) and applies Adam: no second launch, no atomics
    __threadfence();
    __syncthreads();
    bool bar_failed = false;
    if (tid == 0) {
      if (atomicAdd(p.bar, 1u) == gridDim.x - 1) {
        p.bar[0] = 0;
        __threadfence();
        atomicAdd(p.bar + 1, 1u);
      } else {
        unsigned gnow = bar_gen;
        for (int spin = 0; spin < (1 << 22) && gnow == bar_gen; ++spin) {      // bounded: never hang the device
          __nanosleep(40);
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(gnow) : "l"(p.bar + 1) : "memory");
        }
        if (gnow == bar_gen) { atomicExch(p.bar + 2, 1u); bar_failed = true; }
      }
    }
    mark(18);
    const int t = p.step_state[0] + 1;
    // exchange epoch: the tag / slot selector of the peer-memory all-reduce.  It only ever grows, unlike the Adam step count,
    // which CUDA-graph capture rewinds after its warm-up steps (graph_step._capture): a rewound tag would match the stale
    // packets of the warm-up
    const unsigned ep = (unsigned)p.step_state[2] + 1u;
    float* red = scratch;                           // [4][128] + the two bias-correction scalars + the barrier verdict
    if (tid == 0) {                                 // double-precision pow once per CTA
      const double bc1 = 1.0 - pow(p.beta1, (double)t), bc2 = 1.0 - pow(p.beta2, (double)t);
      red[520] = (float)(p.lr / bc1);
      red[521] = (float)sqrt(bc2);
      red[522] = bar_failed ? 1.f : 0.f;
    }
    __syncthreads();                                // every thread is past the grid barrier from here on
    // a CTA whose barrier wait timed out has no complete gradient to read: it leaves its slice of the parameters and the
    // optimizer state untouched (the host sees the flag before the next checkpoint, FlatAdam.barrier_ok())
    const bool skip_update = red[522] != 0.f;
    if (blockIdx.x == 0 && warp == 1) {             // loss = sum of the per-CTA partials, fixed order
      float l = 0.f;
      for (int k = lane; k < (int)gridDim.x; k += 32) l += __ldcg(p.partials + (size_t)k * PSTRIDE + NGRAD);
      l = warp_sum(l);
      if (lane == 0) p.loss[0] = l;
    }
    const float step_size = red[520], bc2_sqrt = red[521];
    const float omb1 = (float)(1.0 - p.beta1), b2 = (float)p.beta2, omb2 = (float)(1.0 - p.beta2), eps = (float)p.eps;
    const int G = gridDim.x, per = (NGRAD + G - 1) / G, lo = blockIdx.x * per, hi = min(NGRAD, lo + per);
    const int ex = tid & 127, ky = tid >> 7;        // 128 parameters x 4 slices of the partial list per pass
    const bool xchg = p.world > 1;
    for (int base = lo; base < hi && !skip_update; base += 128) {
      const int i = base + ex;
      float s0 = 0.f, gv = 0.f, m0 = 0.f, v0 = 0.f, p0 = 0.f;
      if (i < hi) {
        const float* q = p.partials + i;
        float v[38];
#pragma unroll
        for (int u = 0; u < 38; ++u) v[u] = (ky + 4 * u < G) ? __ldcg(q + (size_t)(ky + 4 * u) * PSTRIDE) : 0.f;
        if (ky == 0) { gv = __ldcg(p.adam_g + i); m0 = __ldcg(p.adam_m + i); v0 = __ldcg(p.adam_v + i); p0 = __ldcg(p.adam_p + i); }
#pragma unroll
        for (int u = 0; u < 38; ++u) s0 += v[u];
      }
      red[ky * 128 + ex] = s0;
      __syncthreads();
      if (ky == 0 && i < hi) {
        gv += (red[ex] + red[128 + ex]) + (red[256 + ex] + red[384 + ex]);
        bool upd = true;
        if (xchg) {
          const size_t slot = (size_t)(ep & 1u) * p.world * p.npad;
          const unsigned long long pkt = ((unsigned long long)ep << 32) | __float_as_uint(gv);
          for (int r = 0; r < p.world; ++r)             // push {value, step} into every rank's buffer (mine included)
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(p.ll_peer[r] + slot + (size_t)p.rank * p.npad + i), "l"(pkt)
                         : "memory");
          const unsigned long long* mine = p.ll_peer[p.rank] + slot + i;
          gv = 0.f;
          bool ok = true;
          for (int r = 0; r < p.world; ++r) {           // rank order on every rank
            unsigned long long got = 0;
            int spin = 0;
            do {
              asm volatile("ld.relaxed.sys.global.u64 %0, [%1];\n" : "=l"(got) : "l"(mine + (size_t)r * p.npad) : "memory");
            } while ((unsigned)(got >> 32) != ep && ++spin < (1 << 24));
            ok = ok && (unsigned)(got >> 32) == ep;
            gv += __uint_as_float((unsigned)got);
          }
          if (!ok) atomicExch(p.bar + 2, 1u);           // a peer never delivered: flagged, never hangs
          upd = ok;
        }
        if (upd) {                                      // incomplete sum: leave this parameter and its moments as they are
          const float mv = m0 + omb1 * (gv - m0);
          const float vv = v0 * b2 + omb2 * gv * gv;
          p.adam_m[i] = mv;
          p.adam_v[i] = vv;
          p.adam_p[i] = p0 - step_size * (mv / (sqrtf(vv) / bc2_sqrt + eps));
        }
        p.adam_g[i] = 0.f;
      }
      __syncthreads();
    }
    mark(19);
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(&p.step_state[1], 1) == (int)gridDim.x - 1) {
        p.step_state[1] = 0;
        p.step_state[0] = t;
        p.step_state[2] = (int)ep;
      }
    }
  }

}  // namespace cf
}  // namespace cgs
