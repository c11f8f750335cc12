/* cgs_b200 — C-ABI of the B200-native Critic/Hourglass conv hot path.
 *
 * The reference has no FFI: its hot path is the torch.nn call graph of
 * nets.py (NewCritic nets.py:160-212, UnetDecoder nets.py:452-523) plus the loss
 * expressions of main.py:191-198 and main.py:364-462.  Each entry point below names
 * the reference ops it replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the
 *     caller (the Python host allocates torch tensors and passes data_ptr());
 *   - activations are NHWC fp32 ([B,H,W,C] contiguous == torch channels_last);
 *     weights keep the reference's OIHW fp32 state_dict layout and are never
 *     repacked in place;
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), no internal
 *     synchronisation, no allocation; safe to capture into a CUDA graph;
 *   - return 0 on success, negative cgs_status on error; cgs_last_error() gives text.
 */
#ifndef CGS_B200_H
#define CGS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum cgs_status { CGS_OK = 0, CGS_EINVAL = -1, CGS_ECUDA = -2, CGS_EUNSUPPORTED = -3 };

/* How a conv operand is produced on the fly while it is staged into shared memory. */
enum cgs_src_mode {
  CGS_SRC_PLAIN = 0,    /* a[B,H,W,C]; optional b = multiplicative mask of the same shape (Dropout, nets.py:179,183) */
  CGS_SRC_CATUP = 1,    /* cat(a[B,H,W,C0], nearest_up(b[B,H>>shift,W>>shift,C-C0])) — T.cat + nn.Upsample, nets.py:503-520 */
  CGS_SRC_POOLBWD = 2,  /* grad of conv output through ReLU+MaxPool2d(2): a=dE[B,H/2,W/2,C], b=E (pooled fwd output), idx=argmax */
  CGS_SRC_SIGGRAD = 3,  /* a=dZ, b=Z: dZ*Z*(1-Z) — Sigmoid backward (nets.py:491) */
  CGS_SRC_LEAKYGRAD = 4,/* a=dOut, b=out: dOut*(out>0?1:slope) — LeakyReLU(0.01) backward (nets.py:462,489) */
  CGS_SRC_U8ROLL = 5    /* a = (const float*)uint8 frames [B,H,W,C]: value = u8[n,y,(x+roll) mod W,c]/255 — the
                           `.float()/255.0` cast and Handler.shift_batch roll (main.py:189,584-591) fused into the
                           operand load; roll = `shift` field, or *(const int32_t*)b when b != NULL */
};

typedef struct cgs_src {
  int32_t mode;
  int32_t C;            /* channels presented to the conv */
  int32_t C0;           /* CATUP: channels taken from a */
  int32_t shift;        /* CATUP: log2 upsample factor of b (1 or 2) */
  const float* a;
  const float* b;
  const uint8_t* idx;   /* POOLBWD: 2-bit window position of the first max, one byte per pooled element */
} cgs_src;

/* Epilogues of the 3x3 conv kernel. */
enum cgs_epi {
  CGS_EPI_LINEAR = 0,     /* out = acc + bias                         (dec convs: no activation, nets.py:500-517) */
  CGS_EPI_LEAKY = 1,      /* out = leaky_relu(acc + bias, 0.01)       (masker[0..1], nets.py:488-489) */
  CGS_EPI_RELU_POOL = 2,  /* out = maxpool2(relu(acc + bias)), idx    (features[k..k+2], nets.py:170-182) */
  CGS_EPI_SIGMOID = 3,    /* out = sigmoid(acc + bias) [+ hard = out >= thresh]  (masker[2..3], main.py:1164) */
  CGS_EPI_MUL = 4,        /* out = acc * mul  (dgrad through Dropout) */
  CGS_EPI_SPLIT_UP = 5    /* channels < C0 -> out[B,H,W,C0]; the rest summed over the upsample window into out2 (cat + Upsample backward) */
};

typedef struct cgs_conv3x3_args {
  cgs_src src;            /* input operand, Cin = src.C */
  const float* w;         /* OIHW [Cout,Cin,3,3]; if transposed: the forward layer's [Cin,Cout,3,3] used as dgrad filter */
  const float* bias;      /* [Cout] or NULL */
  int32_t transposed;     /* 0 = fprop, 1 = dgrad (rotate 180 deg, swap in/out) */
  int32_t B, H, W;        /* conv resolution (input == output, stride 1 pad 1) */
  int32_t Cout;
  int32_t epi;
  float* out;             /* see cgs_epi */
  float* out2;            /* SPLIT_UP: [B,H>>shift2,W>>shift2,Cout-C0], must be zeroed when shift2==2 */
  uint8_t* idx_out;       /* RELU_POOL: argmax bytes (may be NULL); SIGMOID: hard mask bytes (may be NULL) */
  const float* mul;       /* MUL: multiplier, same shape as out */
  int32_t C0;             /* SPLIT_UP: split point */
  int32_t shift2;         /* SPLIT_UP: 1 or 2 */
  float thresh;           /* SIGMOID hard-mask threshold (>=) */
  int32_t precision;      /* CGS_FP32: CUDA-core FFMA, exact fp32 (rtol 1e-4 parity path);
                             CGS_TF32: tcgen05.mma kind::tf32 with fp32 accumulation in TMEM where the shape is
                             covered (H%16==0, W%8==0), fp32 kernel otherwise */
} cgs_conv3x3_args;

enum cgs_precision { CGS_FP32 = 0, CGS_TF32 = 1 };

/* Conv2d(k=3,s=1,p=1) fprop or dgrad with fused prologue/epilogue.
 * Replaces: nn.Conv2d + ReLU + MaxPool2d (nets.py:170-182), T.cat + nn.Upsample + nn.Conv2d
 * (nets.py:503-521), LeakyReLU / Sigmoid (nets.py:489-491) and autograd's conv input-gradient. */
int cgs_conv3x3(const cgs_conv3x3_args* a, void* stream);

/* First encoder stage straight from raw frames: uint8 NHWC [B,H,W,3] -> /255 -> circular W-roll (roll, or *roll_dev) ->
 * Conv2d(3,Cout,3,1,1) + bias -> ReLU -> MaxPool2d(2) -> e0 [B,H/2,W/2,Cout] (+ argmax bytes idx0, may be NULL).
 * Replaces `X.permute(0,3,1,2).float()/255`, Handler.shift_batch and features[0..2] (main.py:185-189, nets.py:170-172).
 * H, W multiples of 16, Cout a multiple of 8. */
int cgs_conv_rgb_fwd(const uint8_t* frames, int32_t B, int32_t H, int32_t W, int32_t roll, const int32_t* roll_dev,
                     const float* w, const float* bias, int32_t Cout, float* e0, uint8_t* idx0, void* stream);

typedef struct cgs_wgrad3x3_args {
  cgs_src x;              /* forward input operand (Cin = x.C) */
  cgs_src dy;             /* gradient w.r.t. the conv output (Cout = dy.C) */
  int32_t B, H, W;
  float* dw;              /* OIHW [Cout,Cin,3,3], ACCUMULATED into (caller zeroes) */
  float* db;              /* [Cout] accumulated into, or NULL */
  int32_t precision;      /* CGS_FP32: FFMA; CGS_TF32: tensor-core mma (TF32 operands, fp32 accumulate) for H,W >= 8 */
} cgs_wgrad3x3_args;

/* Weight + bias gradient of Conv2d(k=3,s=1,p=1): replaces autograd's conv weight-gradient
 * for every conv row of SURVEY.md §8a (a20'). */
int cgs_wgrad3x3(const cgs_wgrad3x3_args* a, void* stream);

/* NewCritic tail: features[13..15] + crit (nets.py:183-195): Dropout, Conv2d(16c,32c,4) on the
 * 4x4 map, ReLU (-> embeds[4]), Flatten, Linear, ReLU, Dropout, Linear, Sigmoid.
 * e3 [B,4,4,C3] NHWC; m_e3 / m_v optional dropout masks ([B,4,4,C3], [B,NB]);
 * outputs e4 [B,NB], v [B,NB] (post-ReLU, pre-dropout, saved for backward), pred [B]. */
int cgs_head_fwd(const float* e3, const float* m_e3, const float* m_v,
                 const float* w14, const float* b14, const float* w1, const float* b1,
                 const float* w2, const float* b2,
                 int32_t B, int32_t C3, int32_t NB,
                 float* e4, float* v, float* pred, void* stream);

/* Backward of cgs_head_fwd.  dpred [B] and optional de4 [B,NB] (gradient reaching embeds[4]
 * from the decoder).  Parameter gradients are ACCUMULATED (any may be NULL to skip all
 * parameter gradients: pass dw14 == NULL); de3 [B,4,4,C3] is written (NULL to skip). */
int cgs_head_bwd(const float* e3, const float* m_e3, const float* m_v,
                 const float* w14, const float* w1, const float* w2,
                 const float* e4, const float* v, const float* pred,
                 const float* dpred, const float* de4,
                 int32_t B, int32_t C3, int32_t NB,
                 float* dw14, float* db14, float* dw1, float* db1, float* dw2, float* db2,
                 float* de3, void* stream);

/* Fused tail of NewCritic: features[9..15] + crit (nets.py:179-195) in one kernel: [Dropout] Conv2d(8c,16c,3,1,1)
 * ReLU MaxPool2d(2) (-> embeds[3], argmax) + everything cgs_head_fwd does.  e2 [B,8,8,C2] NHWC; m_e2 / m_e3 / m_v
 * optional dropout masks.  cgs_tail_supported() tells whether the shapes fit the kernel's shared-memory budget
 * (returns 1/0); otherwise use cgs_conv3x3 + cgs_head_fwd. */
int cgs_tail_supported(int32_t B, int32_t C2, int32_t C3, int32_t NB);
int cgs_tail_fwd(const float* e2, const float* m_e2, const float* m_e3, const float* m_v,
                 const float* w3, const float* b3, const float* w14, const float* b14,
                 const float* w1, const float* b1, const float* w2, const float* b2,
                 int32_t B, int32_t C2, int32_t C3, int32_t NB,
                 float* e3, uint8_t* idx3, float* e4, float* v, float* pred, void* stream);
/* Backward of cgs_tail_fwd.  Incoming gradients dpred [B], de3_ext [B,4,4,C3] (from the decoder skip), de4 [B,NB] are
 * each optional (not all NULL).  Parameter gradients accumulated (all-or-none, dw14 == NULL skips); de2 written or NULL. */
int cgs_tail_bwd(const float* e2, const float* m_e2, const float* m_e3, const float* m_v,
                 const float* w3, const float* w14, const float* w1, const float* w2,
                 const float* e3, const uint8_t* idx3, const float* e4, const float* v, const float* pred,
                 const float* dpred, const float* de3_ext, const float* de4,
                 int32_t B, int32_t C2, int32_t C3, int32_t NB,
                 float* dw3, float* db3, float* dw14, float* db14, float* dw1, float* db1, float* dw2, float* db2,
                 float* de2, void* stream);

/* The 14 NewCritic tensors in state_dict order (nets.py:169-195): features.{0,3,6,10,14}.{weight,bias} OIHW,
 * crit.{1,4}.{weight,bias}; used both for the parameters (read) and for their gradients (accumulated into). */
typedef struct cgs_critic_weights {
  float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4, *wl1, *bl1, *wl2, *bl2;
} cgs_critic_weights;

/* One whole critic_pipe training step minus Adam (main.py:185-198) in ONE kernel, for the chfak=1 geometry
 * (cgs_critic_fused_supported): uint8 NHWC frames [B,64,64,3] -> /255 -> shift_batch roll (roll, or *roll_dev) ->
 * NewCritic.forward in train mode (m_e2 [B,8,8,8], m_e3 [B,4,4,16], m_v [B,32] multiplicative dropout masks, NULL =
 * identity) -> F.mse_loss / F.binary_cross_entropy against target [B] -> backward.  Every activation of a frame stays
 * in shared memory; TF32 tensor-core (mma.sync) convolutions with fp32 accumulation, fp32 head.
 * Outputs: pred [B]; loss[0] = mean loss over the B frames; the gradient of (loss_grad * loss) is ACCUMULATED into *g. */
int cgs_critic_fused_supported(int32_t C0, int32_t C1, int32_t C2, int32_t C3, int32_t NB);
/* Gradient delivery: either REDs into the 14 tensors of *g (partials == NULL), or — cheaper, and bit-reproducible — one
 * partial gradient vector per CTA, partials[cta * cgs_critic_fused_partial_stride() + i], i over the 11,873 critic
 * parameters in state_dict order, cta < cgs_critic_fused_grid(B); the consumer sums them (cgs_adam_step_partials /
 * cgs_reduce_partials).  Every partial row is fully overwritten by each call. */
int cgs_critic_fused_grid(int32_t B);
int cgs_critic_fused_partial_stride(void);
/* Optional in-kernel optimizer of cgs_critic_train_fused (single GPU, the optimizer bucket is exactly the 11,873 critic
 * parameters in state_dict order): after a grid barrier every CTA sums its slice of the partial vectors and applies Adam
 * (torch.optim.Adam defaults semantics, as cgs_adam_step; g is added to the gradient and cleared).
 * barrier: 4 x uint32 device words, zeroed once; word 2 != 0 afterwards = a CTA timed out at the barrier. */
typedef struct cgs_adam_args {
  float *p, *g, *m, *v;
  double lr, beta1, beta2, eps;
  int32_t* step_state;
  uint32_t* barrier;
  /* world > 1: the kernel also all-reduces the gradient over NVLink peer memory before the update (one launch per
   * data-parallel step), low-latency protocol: every rank owns a symmetric receive buffer of 2*world*npad 8-byte words
   * ({fp32 value, step tag}; zeroed once); each CTA pushes its slice of the local gradient into every rank's buffer and
   * polls its own until all tags equal the step, then sums in rank order.
   * peer_recv: HOST array of `world` device addresses (this process's mapping of each rank's receive buffer). */
  int32_t world, rank;
  int64_t npad;
  const uint64_t* peer_recv;
} cgs_adam_args;

/* Dropout: explicit masks (m_e2, m_e3, m_v), or rng_state != NULL: the kernel draws them itself — the SAME Philox stream
 * cgs_dropout_masks(out, B*800, p_drop, seed, rng_state) would have written for shapes [B,8,8,8 | B,4,4,16 | B,32] — and
 * advances rng_state like that call does; or neither (eval / p = 0). */
int cgs_critic_train_fused(const uint8_t* frames, const float* target, int32_t B, int32_t roll, const int32_t* roll_dev,
                           const float* m_e2, const float* m_e3, const float* m_v,
                           float p_drop, uint64_t seed, uint64_t* rng_state,
                           const cgs_critic_weights* w, const cgs_critic_weights* g, float* partials,
                           const cgs_adam_args* adam, float loss_grad, int32_t bce, float* pred, float* loss, void* stream);

/* g[offset + j] += sum_k partials[k * stride + j], j < len, summed in a fixed order (data-parallel path: all-reduce next). */
int cgs_reduce_partials(float* g, int64_t n, const float* partials, int32_t n_partials, int64_t stride,
                        int64_t offset, int64_t len, void* stream);
/* cgs_adam_step whose gradient is g[i] + sum_k partials[k * stride + (i - offset)] (inside [offset, offset + len));
 * g is cleared.  Replaces the gradient-REDs + Adam pair of the single-GPU critic step. */
int cgs_adam_step_partials(float* p, float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                           double eps, int32_t* step_state, float grad_scale, const float* partials,
                           int32_t n_partials, int64_t stride, int64_t offset, int64_t len, void* stream);

/* Frozen-critic loss and its gradient w.r.t. the INPUT frames in ONE kernel (chfak=1 geometry; the XG variant of the
 * whole-step kernel): x [B,64,64,3] fp32 NHWC -> NewCritic.forward (dropout as for cgs_critic_train_fused) ->
 * F.mse_loss / F.binary_cross_entropy against target [B] -> backward through the critic to the frames; no weight
 * gradients.  This is `critic(replaced)` / `critic(injected)` of the Hourglass loop with their losses and the backward
 * into the blend (main.py:396-411), and `pred.mean().backward(); batch.grad` of the saliency baseline (main.py:949-951).
 * Outputs: pred [B]; loss[0] = mean loss; dx [B,64,64,3] = loss_grad * d loss / d x.
 * bce: 0 = MSE, 1 = BCE, 2 = loss = mean(pred), target ignored (saliency baseline: `pred.mean().backward()`, main.py:949).
 * dx == NULL selects the forward-only variant (target and loss may then be NULL): pred only, e.g. `negpred = critic(B)`
 * under no_grad (main.py:365-367) or the predictions of extract_contrastive_data (main.py:238-260). */
int cgs_critic_loss_xgrad(const float* x, const float* target, int32_t B, const float* m_e2, const float* m_e3,
                          const float* m_v, float p_drop, uint64_t seed, uint64_t* rng_state,
                          const cgs_critic_weights* w, float loss_grad, int32_t bce, float* pred, float* loss, float* dx,
                          void* stream);

/* pred [B] = NewCritic.forward(frames) for raw uint8 NHWC frames (rolled by roll / *roll_dev), forward only, ONE kernel:
 * `negpred = critic(B)` under no_grad (main.py:365-367) and the predictions of extract_contrastive_data (main.py:238-260).
 * Dropout as for cgs_critic_train_fused (all NULL / rng_state NULL = eval mode). */
int cgs_critic_forward_frames(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const float* m_e2,
                              const float* m_e3, const float* m_v, float p_drop, uint64_t seed, uint64_t* rng_state,
                              const cgs_critic_weights* w, float* pred, void* stream);

/* `-process` inference, encoder + decoder half, in ONE kernel for the chfak=1 geometry (csrc/infer_fused.cu): uint8 NHWC
 * frames [B,64,64,3] -> /255 -> NewCritic.forward(collect=True) in eval mode (nets.py:197-212) -> UnetDecoder dec[4]..dec[0]
 * with their nearest-upsample + concat operands (nets.py:500-517) -> o0 [B,32,32,8] NHWC (the map `masker` consumes via
 * cat(X, ups(o0)), nets.py:519-521) and pred [B].  TF32 tensor-core convolutions, fp32 accumulate; every intermediate stays
 * in shared memory.  cw: critic tensors; wd4/bd4: dec_model.4; bd3..bd0: dec_model.{3..0}.bias;
 * pack: cgs_infer_pack_floats() floats written by cgs_infer_pack_decoder from dec_model.{3,2,1,0}.weight (OIHW
 * [16,48,3,3] [8,24,3,3] [8,16,3,3] [8,16,3,3]) - re-pack whenever those weights change. */
int cgs_infer_pack_floats(void);
int cgs_infer_pack_decoder(const float* wd3, const float* wd2, const float* wd1, const float* wd0, float* pack, void* stream);
int cgs_infer_fused(const uint8_t* frames, int32_t B, const cgs_critic_weights* cw, const float* wd4, const float* bd4,
                    const float* bd3, const float* bd2, const float* bd1, const float* bd0, const float* pack,
                    float* pred, float* o0, void* stream);

/* The masker half of `-process` in ONE kernel (chfak=1: masker.0 is [16,11,3,3]): cat(X, ups(o0)) -> Conv2d + LeakyReLU(0.01)
 * -> Conv2d(16,1) -> Sigmoid (nets.py:488-491, 519-523) [-> hard = mask >= thresh, main.py:1164].  frames: uint8 NHWC
 * [B,64,64,3]; o0 [B,32,32,8] NHWC fp32 as written by cgs_infer_fused; mask [B,64,64] fp32; hard [B,64,64] uint8 or NULL.
 * The 16-channel 64x64 intermediate only ever exists as an 18-row band in shared memory. */
int cgs_masker_fused(const uint8_t* frames, const float* o0, int32_t B, const float* wm0, const float* bm0,
                     const float* wm2, const float* bm2, float thresh, float* mask, uint8_t* hard, void* stream);

/* Data-parallel gradient exchange fused with Adam over NVLink peer memory (csrc/p2p_adam.cu).  Every rank owns a SYMMETRIC
 * gradient buffer sym[2][npad] and a flag pad (>= 16 uint32, zeroed once), both mapped into all peers.
 * cgs_p2p_stage: sym_local[slot][i] = g[i] + sum_k partials[k*stride + i - offset] (n_partials may be 0), g cleared;
 *   slot = (step_state[2] + 1) & 1 is read on the device, so the call is CUDA-graph replayable.  step_state here (and in
 *   cgs_adam_args of the whole-step kernel) is 4 x int32 {Adam steps applied, ticket, exchange epoch, -}: the exchange epoch
 *   selects the slot and tags the announcements; it only ever grows, whereas a caller may rewind the Adam step count.
 * cgs_p2p_allreduce_adam: announce step t to all peers, wait for theirs (bounded spin; *err_flag = 1 on time-out), sum the
 *   `world` buffers in rank order through peer loads, apply Adam (torch defaults, as cgs_adam_step) to p/m/v.
 *   peer_bufs / peer_flags: HOST arrays of `world` device addresses (this process's mappings of each rank's allocation).
 * Replaces torch.distributed.all_reduce(gradient bucket) + optimizer.step() of the data-parallel loops. */
int cgs_p2p_stage(float* g, int64_t n, int64_t npad, float* sym_local, const float* partials, int32_t n_partials,
                  int64_t stride, int64_t offset, int64_t len, const int32_t* step_state, void* stream);
int cgs_p2p_allreduce_adam(float* p, float* m, float* v, int64_t n, int64_t npad, const uint64_t* peer_bufs,
                           const uint64_t* peer_flags, int32_t rank, int32_t world, double lr, double beta1, double beta2,
                           double eps, int32_t* step_state, float grad_scale, int32_t* err_flag, void* stream);

/* The two scored blends of one Hourglass step in ONE kernel (chfak=1 geometry; MODE 3 of the whole-step kernel),
 * replacing `replaced = A*(1-Z)+Z*B; F.mse_loss(critic(replaced), negpred)`, `injected = B*(1-Z)+Z*A;
 * F.mse_loss(critic(injected), pred)` and the mask regulariser with their backward into Z (reference main.py:395-429):
 * frames_a (rolled by `roll` / *roll_dev like shift_batch, main.py:355-357) and frames_b are uint8 [B,64,64,3], z is the
 * mask [B,64,64]; the blends are formed in shared memory, scored by the FROZEN critic (dropout: forced masks for the
 * replace and inject passes, or the module's Philox stream - two consecutive calls), and the input gradient is
 * contracted with (B - A) resp. (A - B) inside the kernel.  target_inject == NULL skips the inject pass (-noinject).
 * Outputs: pred_replace / pred_inject [B]; losses[4] = {replace, inject, L1 term, L2 term} (means, as the reference
 * logs them); dz [B,64,64] = loss_grad * d(replace + inject + L1 + L2)/dZ.  vpred != NULL: non-static regulariser
 * weight 1 - vpred[n] (main.py:418). */
int cgs_hg_score(const uint8_t* frames_a, const uint8_t* frames_b, int32_t B, int32_t roll, const int32_t* roll_dev,
                 const float* z, const float* target_replace, const float* target_inject,
                 const float* m_e2, const float* m_e3, const float* m_v,
                 const float* m_e2_inj, const float* m_e3_inj, const float* m_v_inj,
                 float p_drop, uint64_t seed, uint64_t* rng_state, const cgs_critic_weights* w, float loss_grad,
                 const float* vpred, float l1, float l2, float* pred_replace, float* pred_inject, float* losses,
                 float* dz, void* stream);

/* Debug only: clock64() phase trace of CTA 0 of the fused critic kernel into dev_buf[4*24] (NULL disables). */
int cgs_critic_fused_set_trace(long long* dev_buf);

/* Dense layer out[B,N] = in[B,K] * w[N,K]^T + bias: UnetDecoder.dec[4], the 1x1 conv on the
 * 1x1 bottleneck (nets.py:484,500-501). */
int cgs_dense_fwd(const float* in, const float* w, const float* bias,
                  int32_t B, int32_t K, int32_t N, float* out, void* stream);
/* Its backward: din[B,K] written (or NULL); dw[N,K], db[N] accumulated (or NULL). */
int cgs_dense_bwd(const float* in, const float* w, const float* dout,
                  int32_t B, int32_t K, int32_t N, float* din, float* dw, float* db, void* stream);

/* Occlusion blend out = a*(1-z) + z*b over [B,H,W,C] with z [B,H,W,1] (main.py:395,406). */
int cgs_occlude_fwd(const float* a, const float* b, const float* z, int64_t npix, int32_t C,
                    float* out, void* stream);
/* dz[p] = sum_c (b-a)[p,c]*g[p,c]; optional da = g*(1-z), db_ = g*z. */
int cgs_occlude_bwd(const float* a, const float* b, const float* z, const float* g,
                    int64_t npix, int32_t C, float* dz, float* da, float* db_, void* stream);

/* loss[0] = scale * mean((p-t)^2) over n (F.mse_loss, main.py:195,400,411); grad (optional)
 * = gscale * 2*(p-t)/n.  `bce` != 0 selects F.binary_cross_entropy (main.py:193). */
int cgs_pred_loss(const float* p, const float* t, int32_t n, int32_t bce, float gscale,
                  float* loss, float* grad, void* stream);

/* Mask regulariser (main.py:415-429): loss[0] = L1*mean|vf*z| + L2*mean((vf*z)^2) with
 * vf = 1 (staticnorm) or 1 - vpred[frame]; grad (optional) [n] written = gscale * d loss / d z. */
int cgs_mask_reg(const float* z, const float* vpred, int64_t n, int32_t per_frame,
                 float l1, float l2, float gscale, float* loss, float* grad, void* stream);

/* uint8 NHWC frames -> fp32 /255 with the circular W-roll of Handler.shift_batch
 * (main.py:584-591, 189): out[n,y,x,c] = in[n,y,(x+roll) mod W,c]/255.  If roll_dev != NULL the
 * roll is read from device memory instead (so a captured CUDA graph can vary it per replay). */
int cgs_frames_to_float(const uint8_t* in, int32_t B, int32_t H, int32_t W, int32_t C,
                        int32_t roll, const int32_t* roll_dev, float* out, void* stream);

/* Adam, torch.optim.Adam defaults (main.py:178,331-334), over a flat parameter bucket.  step_state = {number of steps
 * already applied, ticket} in device memory: the kernel applies step state[0]+1 and publishes it itself, so the call can
 * sit inside a CUDA graph with no counter-increment launch.  clear_grad != 0 zeroes g after use (fused zero_grad). */
int cgs_adam_step(float* p, float* g, float* m, float* v, int64_t n,
                  double lr, double beta1, double beta2, double eps, int32_t* step_state,
                  float grad_scale, int32_t clear_grad, void* stream);

/* ---- bf16 whole-frame Hourglass kernels (csrc/hg_forward.cu, csrc/hg_backward.cu), chfak=1 geometry ----------------------
 * UnetDecoder parameters in state_dict order (reference nets.py:479-491): dec_model.{0..4}, masker.{0,2}; OIHW fp32. */
typedef struct cgs_masker_weights {
  float *wd0, *bd0, *wd1, *bd1, *wd2, *bd2, *wd3, *bd3, *wd4, *bd4, *wm0, *bm0, *wm2, *bm2;
} cgs_masker_weights;

/* Weight fragments of every convolution of both networks in bf16 mma.sync B-fragment order: cgs_hg_pack_words() uint32 words,
 * 16-byte aligned; re-pack whenever a weight changes (one small launch). */
int cgs_hg_pack_words(void);
int cgs_hg_pack(const cgs_critic_weights* cw, const cgs_masker_weights* mw, uint32_t* pack, void* stream);
/* The Hourglass forward in ONE kernel: uint8 NHWC frames [B,64,64,3] (circularly rolled along W by roll / *roll_dev, the
 * shift_batch augmentation, main.py:584-591) -> /255 -> NewCritic.forward(collect=True) (nets.py:197-212) ->
 * UnetDecoder.forward (nets.py:494-523) -> z [B,64,64] fp32 mask, pred [B], hard [B,64,64] = z >= thresh (main.py:1164;
 * may be NULL).  train != 0: dropout applied (forced masks m_e2 [B,8,8,8] / m_e3 [B,4,4,16] / m_v [B,32], or drawn in the
 * kernel from (seed, rng_state) exactly as cgs_dropout_masks would).  tape != NULL: B * cgs_hg_tape_bytes() bytes receive
 * every skip / decoder activation of each frame (haloed bf16 planes) for cgs_hg_backward.
 * Replaces `critic(A, collect=True)` + `masker(A, embeds)` of main.py:364,391 and main.py:1139-1164. */
int cgs_hg_tape_bytes(void);
int cgs_hg_forward(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const cgs_critic_weights* cw,
                   const cgs_masker_weights* mw, const uint32_t* pack, int32_t train, const float* m_e2, const float* m_e3,
                   const float* m_v, float p_drop, uint64_t seed, uint64_t* rng_state, float thresh, float* pred, float* z,
                   uint8_t* hard, void* tape, void* stream);
/* The masker's whole backward in ONE kernel (the autograd backward of nets.py:494-523 w.r.t. the 13,785 UnetDecoder
 * parameters, critic frozen: main.py:334, 462): frames / roll / tape / z as given to / written by cgs_hg_forward, dz
 * [B,64,64] = d loss / d z.  Output: one partial gradient vector per CTA, partials[cta * cgs_hg_partial_stride() + i],
 * i in state_dict order, cta < cgs_hg_grid(B); cgs_adam_step_partials / cgs_reduce_partials sum them in a fixed order.
 * debug: NULL, or cgs_hg_debug_floats() floats that receive frame 0's decoder-activation gradients (tests only). */
int cgs_hg_grid(int32_t B);
int cgs_hg_partial_stride(void);
int cgs_hg_debug_floats(void);
/* Non-zero if a bounded mbarrier wait (TMA bulk load) of cgs_hg_backward ever timed out; synchronises. */
int cgs_hg_status(void);
/* Debug only: clock64() phase traces of CTA 0 of cgs_hg_forward / cgs_hg_backward / cgs_hg_score_bf16 into 64 int64 each
 * (NULL disables). */
int cgs_hg_set_trace(long long* fwd_buf, long long* bwd_buf, long long* score_buf);
int cgs_hg_backward(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const cgs_masker_weights* mw,
                    const uint32_t* pack, const void* tape, const float* z, const float* dz, float* partials, float* debug,
                    void* stream);

/* The critic-scoring part of one frozen-critic segmentation_training step in ONE bf16 kernel (csrc/hg_score.cu; main.py:365-367,
 * 395-429): [negpred = critic(B) when target_replace == NULL], replaced = A(1-Z)+ZB and injected = B(1-Z)+ZA scored by the
 * critic against negpred / target_inject (= pred of critic(A)), the L1 / L2 mask regulariser, and dz [B,64,64] = loss_grad *
 * d(sum of the terms) / dZ.  frames_a (rolled by roll / *roll_dev) / frames_b: uint8 NHWC; z [B,64,64]; pack: cgs_hg_pack.
 * masks9: NULL, or a HOST array of 9 device pointers {m_e2, m_e3, m_v} x {critic(B), critic(replaced), critic(injected)} of
 * forced dropout masks (entries of passes that do not run may be NULL); rng_state: masks drawn in the kernel, one call index
 * per pass in that order.  losses[4] = replace, inject, L1, L2 terms.  The TF32 variant is cgs_hg_score. */
int cgs_hg_score_bf16(const uint8_t* frames_a, const uint8_t* frames_b, int32_t B, int32_t roll, const int32_t* roll_dev,
                      const float* z, const float* target_replace, const float* target_inject, const float* const* masks9,
                      float p_drop, uint64_t seed, uint64_t* rng_state, const cgs_critic_weights* w, const uint32_t* pack,
                      float loss_grad, const float* vpred, float l1, float l2, float* negpred, float* pred_replace,
                      float* pred_inject, float* losses, float* dz, void* stream);

/* The bf16 variant of cgs_critic_train_fused (csrc/hg_critic.cu): same arguments (minus the RED gradient struct: the gradient
 * leaves as per-CTA partial vectors only), same outputs, same optional in-kernel Adam / peer-memory all-reduce.  bf16 tensor-core
 * operands in the four 3x3 convolutions (forward, input and weight gradients), fp32 accumulation, fp32 head. */
int cgs_critic_train_bf16(const uint8_t* frames, const float* target, int32_t B, int32_t roll, const int32_t* roll_dev,
                          const float* m_e2, const float* m_e3, const float* m_v, float p_drop, uint64_t seed, uint64_t* rng_state,
                          const cgs_critic_weights* w, float* partials, const cgs_adam_args* adam, float loss_grad, int32_t bce,
                          float* pred, float* loss, void* stream);
/* Debug only: clock64() phase trace of CTA 0 of cgs_critic_train_bf16 into 64 int64 (NULL disables). */
int cgs_hg_set_trace_critic(long long* buf);

/* ---- the wide (chfak > 1) path: TMA-fed tcgen05 / TMEM convolutions, bf16 operands, fp32 accumulation (csrc/wide_tc.cu) ----
 * Activations are bf16 "chunk-planar" [B][C/8][H][W][8]; filters stay fp32 OIHW (packed to bf16 operand tiles in the kernel).
 * Replaces nn.Conv2d(k=3, padding=1) + ReLU + MaxPool2d(2) [+ Dropout] of nets.py:170-183 and their autograd backward. */
enum {
  CGS_WIDE_EPI_PLAIN = 0,     /* out[B][Cout/8][H][W][8] = conv + bias */
  CGS_WIDE_EPI_RELU_POOL = 1, /* out[B][Cout/8][H/2][W/2][8] = maxpool2(relu(conv + bias)) [* mask]; idx_out = first-max position
                                 0..3, or 4 where the pooled value is not > 0 (ReLU dead); out_f32: optional NCHW fp32 copy */
  CGS_WIDE_EPI_UNPOOL = 2     /* the conv result is the gradient of a pooled map: out[B][Cout/8][2H][2W][8] receives it [* mask] at
                                 the window position idx_in names and zeros elsewhere (MaxPool + ReLU + Dropout backward) */
};
/* transposed != 0: the input-gradient convolution of a layer with weight w[Cin][Cout][3][3] (x = the output gradient with Cin
 * channels; filters rotated, in/out swapped).  mask: NHWC fp32 [B][h][w][Cout] at the resolution of `out`'s pooled map
 * (RELU_POOL) or of the conv itself (UNPOOL), or NULL.
 * wpacked: the filters as bf16 operand tiles written by cgs_wide_pack (then w may be NULL); NULL: every CTA packs them from w. */
int cgs_wide_conv3x3(const void* x, int32_t B, int32_t H, int32_t W, int32_t Cin, const float* w, const void* wpacked, const float* bias,
                     int32_t Cout, int32_t transposed, int32_t epi, void* out, float* out_f32, uint8_t* idx_out, const uint8_t* idx_in,
                     const float* mask, void* stream);
/* fp32 OIHW filters -> the bf16 operand tiles cgs_wide_conv3x3 keeps in shared memory, up to 8 (layer, direction) jobs in one
 * launch (the filters change with every optimizer step); out: cgs_wide_packed_bytes(Cin, Cout) bytes, 16-byte aligned. */
typedef struct {
  const float* w;
  void* out;
  int32_t Cin, Cout, transposed; /* GEMM view, as in cgs_wide_conv3x3 */
} cgs_wide_packjob;
int cgs_wide_pack(const cgs_wide_packjob* jobs, int32_t njobs, void* stream);
int64_t cgs_wide_packed_bytes(int32_t Cin, int32_t Cout);
/* dw[Cout][Cin][3][3] += sum_pixels dy (x) shifted x, db[Cout] += sum_pixels dy (db may be NULL): x [B][Cin/8][H][W][8],
 * dy [B][Cout/8][H][W][8] bf16; GEMM with K = pixels on tcgen05, per-CTA partial tiles summed in fixed order by a second
 * kernel.  workspace: cgs_wide_wgrad_workspace(B, H, W, Cout) floats.  Cin <= 40. */
int cgs_wide_wgrad3x3(const void* x, const void* dy, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, float* dw, float* db,
                      float* workspace, int64_t workspace_floats, void* stream);
int64_t cgs_wide_wgrad_workspace(int32_t B, int32_t H, int32_t W, int32_t Cout);
/* features.0 of the wide path (3 -> C0 on 64x64, bf16 mma.sync on the pair-duplicated frame): uint8 frames [B,64,64,3] -> /255 ->
 * roll -> conv + ReLU + pool -> e0 [B][C0/8][32][32][8] bf16 + arg-max bytes; and its weight / bias gradient from the pooled
 * gradient de0 (same layout) + arg-max bytes: dw0 [C0][3][3][3] +=, db0 [C0] += (workspace: 148 * 48 * C0 floats). */
int cgs_wide_conv0_fwd(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const float* w0, const float* b0,
                       int32_t C0, void* e0, uint8_t* idx0, void* stream);
int cgs_wide_conv0_wgrad(const uint8_t* frames, int32_t B, int32_t roll, const int32_t* roll_dev, const void* de0, const uint8_t* idx0,
                         int32_t C0, float* dw0, float* db0, float* workspace, int64_t workspace_floats, void* stream);
/* C[M][N] (+)= op(A) op(B) in TF32 (fp32 accumulate): A[m*lda + k] (a_k_contiguous) or A[k*lda + m]; B[n*ldb + k]
 * (b_k_contiguous) or B[k*ldb + n]; then + bias[n], zero where gate[m*ldc + n] <= 0, ReLU, accumulate.  The head's GEMMs:
 * features.14 (a 4x4 valid conv on a 4x4 map, nets.py:186), crit.1 (nets.py:190) and their input / weight gradients.
 * splits > 1: K is cut over gridDim.z; partial tiles meet in ws [splits][M][N] and the last CTA of a tile sums them in order
 * (deterministic); counters: >= tiles ints, zero before the first call (the kernel leaves them zero). */
int cgs_wide_gemm(const float* A, int32_t a_k_contiguous, int32_t lda, const float* Bm, int32_t b_k_contiguous, int32_t ldb, float* Cm,
                  int32_t ldc, int32_t M, int32_t N, int32_t K, const float* bias, const float* gate, int32_t relu, int32_t accumulate,
                  int32_t splits, float* ws, int32_t* counters, void* stream);
/* crit.3 Dropout (mv or NULL), crit.4 Linear(nb, 1), Sigmoid, MSE / BCE (main.py:192-195) on V = crit.2's output [B][nb], one
 * warp per frame: pred [B], lterm [B] (the frame's loss term), dz [B], dV = d loss / d V with the loss scaled by loss_grad / B,
 * U = dz * V * mv (its column sum is crit.4's weight gradient). */
int cgs_wide_head_mid(const float* V, const float* mv, const float* wl2, const float* bl2, const float* target, int32_t B, int32_t nb,
                      float loss_grad, int32_t bce, float* pred, float* lterm, float* dV, float* dz, float* U, void* stream);
/* Up to 5 column sums in one launch: out[j] = (accumulate ? out[j] : 0) + scale * sum_b X[b*n + j], fixed order (bias gradients,
 * crit.4's weight gradient, the mean loss). */
typedef struct {
  const float* X;
  float* out;
  int32_t n;
  float scale;
  int32_t accumulate;
} cgs_wide_coljob;
int cgs_wide_colsums(const cgs_wide_coljob* jobs, int32_t njobs, int32_t B, void* stream);
/* Gradient of features.13's output de3 [B][C3][4][4] fp32 -> Dropout (m3 NHWC or NULL) + MaxPool + ReLU backward (idx3 planar
 * arg-max bytes) -> dy3 [B][C3/8][8][8][8] bf16, the output gradient of features.10. */
int cgs_wide_unpool3(const float* de3, const uint8_t* idx3, const float* m3, int32_t B, int32_t C3, void* dy3, void* stream);
/* Non-zero if a wide kernel ever gave up on an mbarrier (reads a device flag; synchronises). */
int cgs_wide_status(void);
/* Debug only: clock64() marks of CTA 0 of the conv kernel for its first 8 tiles into 64 int64: MMA thread (loop top, TMEM buffer free,
 * operands landed, MMAs issued), first epilogue warp (wait start, accumulator ready, tile stored); [63] = start. */
int cgs_wide_set_trace(long long* dev_buf);

/* ---- formats either side of the path, csrc/edges.cu ---------------------------------------------------
 * out[i] = dataset[idx[i]] for uint8 NHWC frames (12288 bytes each): `Xpos[Hidx]`, `Xneg[Lidx]`, `Xneg[Cidx]` of
 * main.py:345-353 over a device-resident dataset; indices are clamped to [0, nframes). */
int cgs_gather_frames(const uint8_t* dataset, int64_t nframes, const int32_t* idx, int32_t n, uint8_t* out, void* stream);
/* `-process` image outputs (main.py:1212-1223) from mask [B,64,64] fp32 and hard [B,64,64] uint8: raw-mask =
 * (M*255).astype(uint8) and thresholded-mask = hardM*255, replicated to 3 channels: raw, thresholded [B,64,64,3]; or, when
 * concatenated != 0, ONE strip per frame raw[B,64,192,3] = frame | raw-mask | thresholded-mask, with the frame bytes mapped
 * through lut[256] = (b / 255.0 * 255).astype(uint8) (the reference's float64 round trip, not the identity). */
int cgs_mask_images(const float* mask, const uint8_t* hard, int32_t B, const uint8_t* frames, const uint8_t* lut,
                    int32_t concatenated, uint8_t* raw, uint8_t* thresholded, void* stream);
/* Saliency-baseline normalisation (main.py:974-993): sal [B,64,64] >= 0 (`batch.grad.abs().sum(1)`), pred [B];
 * norm = the k-th smallest value of each frame (k = int(4096*thresh)), or *global_norm when non-NULL (-salglobal);
 * out = min(sal / norm * pred, 1), hard = out > thresh; norm_out [B] optional. */
int cgs_saliency_normalize(const float* sal, const float* pred, int32_t B, int32_t k, float thresh, const float* global_norm,
                           float* out, uint8_t* hard, float* norm_out, void* stream);

/* All nn.Dropout masks of one NewCritic forward (nets.py:179,183,192) in one launch: out[i] = Bernoulli(1-p)/(1-p),
 * Philox4x32-10 keyed by (seed, state[0]); state = {call counter, ticket} in device memory, advanced by the kernel
 * itself so that CUDA-graph replays draw fresh masks.  out must be 16-byte aligned. */
int cgs_dropout_masks(float* out, int64_t n, float p, uint64_t seed, uint64_t* state, void* stream);

/* hard[i] = z[i] >= thresh (main.py:1164) or z[i] > thresh when strict (main.py:964). */
int cgs_threshold(const float* z, int64_t n, float thresh, int32_t strict, uint8_t* hard, void* stream);

/* `-eval` (main.py:964, 1265-1270): counts[0] += #(hard & gt), counts[1] += #(hard | gt) with hard = z > thresh (strict,
 * the -eval convention) or z >= thresh; gt = ground-truth mask bytes (non-zero = object).  counts: 2 x uint64 device words the
 * caller zeroes; IoU = counts[0] / counts[1] over everything accumulated.  Exact integer arithmetic. */
int cgs_iou_counts(const float* z, const uint8_t* gt, int64_t n, float thresh, int32_t strict, uint64_t* counts, void* stream);

/* Non-zero if a tcgen05 kernel ever timed out on its completion barrier (reads a device flag; synchronises). */
int cgs_tc_status(void);

/* Debug only: clock64() phase trace of CTA 0 of the tcgen05 conv kernel into dev_buf[16*8] (NULL disables). */
int cgs_tc_set_trace(long long* dev_buf);

const char* cgs_last_error(void);
int cgs_version(void);

#ifdef __cplusplus
}
#endif
#endif
