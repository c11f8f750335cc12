"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU fp32 (optionally fp64) restatement, in functional torch, of the reference's
hot path: the Critic (`NewCritic`) and Hourglass (`UnetDecoder`) forward passes
and the loss terms of the two training loops.  Backward comes from torch autograd
over these functions, exactly as `loss.backward()` does in the reference.

Where the arithmetic lives: the reference delegates every op to PyTorch
(pinned pytorch=1.4.0 cpu, reference requirements.txt:79); the op semantics used
here (conv2d, relu, max_pool2d first-max tie rule, nearest upsample, sigmoid,
mse/l1 mean reductions) are unchanged in the installed torch 2.11.

Pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against outputs of the reference classes themselves,
imported from /root/reference in the authoring container by
tests/golden/make_golden.py; the resulting fixtures are committed under
tests/golden/ and checked by tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import torch
import torch.nn.functional as F


def _sd(sd, dtype):
    return {k: torch.as_tensor(v).to(dtype) for k, v in sd.items()}


# ---- operand-precision models of the tensor-core kernels (tests only) ---------------------------------------------
# The tensor-core paths compute every 3x3 convolution of the reference on operands rounded to TF32 (critic whole-frame
# kernels) or bf16 (Hourglass whole-frame kernels) with fp32 accumulation; everything else (biases, activations,
# pooling, the 4x4 / 1x1 / Linear layers, losses) is fp32.  `q=quant_tf32 | quant_bf16` below reproduces exactly that
# rounding at the same places, with a straight-through gradient, so that the ReLU / max-pool / LeakyReLU *decisions* of the
# model and of the kernel agree and the remaining difference is accumulation order.  q=None is the reference arithmetic.
def quant_bf16(t):
    """round-to-nearest-even to bfloat16 (what __float2bfloat16_rn does), straight-through gradient"""
    return t + (t.detach().bfloat16().float() - t.detach())


def quant_tf32(t):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits; straight-through gradient"""
    d = t.detach()
    r = ((d.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return t + (r - d)


def _conv3(h, w, b, q):
    if q is not None:
        h, w = q(h), q(w)
    return F.conv2d(h, w, b, stride=1, padding=1)


def critic_forward(sd, x, collect=False, masks=None, q=None):
    """NewCritic.forward (reference nets.py:197-212) over the layer list built at
    nets.py:169-195.  `sd` uses the reference state_dict keys.  `masks` is None
    (eval mode: dropout is identity) or a 3-tuple of multiplicative dropout masks
    (values 0 or 1/(1-p)) for features.9 [B,8c,8,8], features.13 [B,16c,4,4] and
    crit.3 [B,32c] (train mode)."""
    embeds = []
    h = x
    for i, key in enumerate(("features.0", "features.3", "features.6", "features.10")):
        h = _conv3(h, sd[key + ".weight"], sd[key + ".bias"], q)                       # nets.py:170,173,176,180
        h = F.relu(h)
        h = F.max_pool2d(h, 2)
        embeds.append(h)                     # nets.py:202-203: collected after each pool
        if i == 2 and masks is not None:
            h = h * masks[0]                 # features.9 Dropout, nets.py:179
        if i == 3 and masks is not None:
            h = h * masks[1]                 # features.13 Dropout, nets.py:183
    h = F.relu(F.conv2d(h, sd["features.14.weight"], sd["features.14.bias"]))          # nets.py:184-185
    embeds.append(h)                         # nets.py:204-205
    v = h.flatten(1)                                                                    # nets.py:189
    v = F.relu(F.linear(v, sd["crit.1.weight"], sd["crit.1.bias"]))                      # nets.py:190-191
    if masks is not None:
        v = v * masks[2]                                                                # nets.py:192
    pred = torch.sigmoid(F.linear(v, sd["crit.4.weight"], sd["crit.4.bias"]))            # nets.py:193-194
    return (pred, embeds) if collect else pred


def decoder_forward(sd, x, embeds, q=None):
    """UnetDecoder.forward (reference nets.py:494-523): no activation between the
    dec convs; every cat is (skip, upsampled); last cat is (image, upsampled)."""
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")                    # nets.py:463
    o = F.conv2d(embeds[4], sd["dec_model.4.weight"], sd["dec_model.4.bias"])           # nets.py:500-501
    o = up(up(o))                                                                       # nets.py:503
    o = _conv3(torch.cat((embeds[3], o), 1), sd["dec_model.3.weight"], sd["dec_model.3.bias"], q)
    o = up(o)
    o = _conv3(torch.cat((embeds[2], o), 1), sd["dec_model.2.weight"], sd["dec_model.2.bias"], q)
    o = up(o)
    o = _conv3(torch.cat((embeds[1], o), 1), sd["dec_model.1.weight"], sd["dec_model.1.bias"], q)
    o = up(o)
    o = _conv3(torch.cat((embeds[0], o), 1), sd["dec_model.0.weight"], sd["dec_model.0.bias"], q)
    o = up(o)
    m = _conv3(torch.cat((x, o), 1), sd["masker.0.weight"], sd["masker.0.bias"], q)             # nets.py:488,519-521
    m = F.leaky_relu(m, 0.01)                                                           # nets.py:462,489
    m = _conv3(m, sd["masker.2.weight"], sd["masker.2.bias"], q)                        # nets.py:490
    return torch.sigmoid(m)                                                             # nets.py:491


def critic_loss(csd, x, y, masks=None, threshrew=False):
    """Loss of one critic_pipe step (reference main.py:191-195)."""
    pred = critic_forward(csd, x, masks=masks).squeeze()
    if threshrew:
        return F.binary_cross_entropy(pred, y), pred
    return F.mse_loss(pred, y), pred


def hourglass_losses(csd, msd, A, Bf, Y=None, live=False, inject=True, L1=0.5, L2=0.0,
                     lfak=5, staticnorm=True, masks=(None, None, None, None), sepsd=None, q_embed=None, q_score=None,
                     q_mask=None):
    """Loss terms of one segmentation_training step (reference main.py:364-429).
    `masks` = dropout masks for the four critic passes in call order
    critic(A), critic(B), critic(replaced), critic(injected).  Returns
    (total, dict of terms, Z)."""
    pred, embeds = critic_forward(csd, A, collect=True, masks=masks[0], q=q_embed)      # main.py:364
    negpred = critic_forward(csd, Bf, masks=masks[1], q=q_score).squeeze().detach()     # main.py:365-367
    pred = pred.squeeze()
    terms = {}
    loss = 0
    if live:
        terms["critic"] = F.mse_loss(pred, Y)                                # main.py:383
        loss = loss + lfak * terms["critic"]
    if sepsd is not None:
        _, embeds = critic_forward(sepsd, A, collect=True, masks=masks[0])   # main.py:389-390
    Z = decoder_forward(msd, A, embeds, q=q_mask)                            # main.py:391
    replaced = A * (1 - Z) + Z * Bf                                          # main.py:395
    terms["replace"] = F.mse_loss(critic_forward(csd, replaced, masks=masks[2], q=q_score).squeeze(), negpred)  # main.py:396-400
    loss = loss + terms["replace"]
    if inject:
        injected = Bf * (1 - Z) + Z * A                                      # main.py:406
        terms["inject"] = F.mse_loss(critic_forward(csd, injected, masks=masks[3], q=q_score).squeeze(), pred.detach())  # main.py:407-411
        loss = loss + terms["inject"]
    vf = 1 if staticnorm else 1 - pred.detach().view(-1, 1, 1, 1)            # main.py:415-419
    if L1:
        terms["L1"] = L1 * F.l1_loss(vf * Z, torch.zeros_like(Z))            # main.py:422
        loss = loss + terms["L1"]
    if L2:
        terms["L2"] = L2 * F.mse_loss(vf * Z, torch.zeros_like(Z))           # main.py:427
        loss = loss + terms["L2"]
    return loss, terms, Z


def segment_batch(csd, msd, batch, threshold):
    """Loop body of Handler.segment (reference main.py:1139-1151, 1164)."""
    pred, embeds = critic_forward(csd, batch, collect=True)
    mask = decoder_forward(msd, batch, embeds)
    return pred, mask, mask >= threshold


def shift_batch(X, xshift, left):
    """Handler.shift_batch (reference main.py:584-591) with the two random draws
    made explicit: circular roll along W (dim 2 of NHWC uint8)."""
    if left:
        return torch.cat((X[:, :, xshift:], X[:, :, :xshift]), dim=2)
    return torch.cat((X[:, :, -xshift:], X[:, :, :-xshift]), dim=2)


def to_input(X_uint8_nhwc):
    """`X.permute(0,3,1,2).float()/255.0` (reference main.py:189, 360): logical NCHW,
    channels_last memory."""
    return torch.as_tensor(X_uint8_nhwc).permute(0, 3, 1, 2).float() / 255.0


def adam_step(params, grads, state, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (reference main.py:178, 331-334), restated so the
    product's flat Adam kernel can be checked without torch.optim."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for i, (p, g) in enumerate(zip(params, grads)):
        m = state.setdefault(("m", i), torch.zeros_like(p))
        v = state.setdefault(("v", i), torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / (1 - b2 ** t) ** 0.5).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))
