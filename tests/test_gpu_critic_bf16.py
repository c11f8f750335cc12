"""GPU: the bf16 whole-step critic kernel (cgs_critic_train_bf16, csrc/hg_critic.cu) against the CPU oracle.

One launch = uint8 frames -> /255 -> shift roll -> NewCritic forward (dropout masks) -> MSE/BCE -> backward (reference
main.py:185-198) with bf16 tensor-core operands in the four 3x3 convolutions (forward, input and weight gradients), fp32
accumulation and an fp32 head.  Two levels, as in tests/test_gpu_hg.py:
  (1) against the oracle evaluated at the kernel's operand precision (`torch_ref.quant_bf16` on the conv operands): the
      ReLU / arg-max decisions are the same, what is left is the bf16 rounding of the back-propagated gradients;
  (2) against the reference arithmetic: what bf16 operands cost (north_star: bf16 path |mask| <= 2e-2; the critic's own
      prediction is held to 1e-2 here)."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import torch_ref
import cgs_b200.synth as synth
from test_gpu_fused import _case, _critic, _oracle, _rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture()
def ops():
    import cgs_b200.ops as o
    o.set_precision("tf32")
    yield o
    o.set_precision("fp32")


def _dev_masks(masks):
    m2, m3, mv = (torch.from_numpy(m).to(DEV) for m in masks)
    return (m2.permute(0, 2, 3, 1).contiguous(), m3.permute(0, 2, 3, 1).contiguous(), mv.contiguous())


def _step(ops, c, Xd, yd, roll, dm, **kw):
    """One bf16 step with the gradient handed to a FlatAdam bucket; returns (loss, pred, flat gradient)."""
    from cgs_b200.train_handler import FlatAdam
    opt = FlatAdam(c.parameters())
    opt.zero_grad()
    loss, pred = ops.critic_train_fused(c, Xd, yd, roll, dm, bf16=True, **kw)
    assert opt.pending_partials is not None
    opt.flush_partials()
    torch.cuda.synchronize()
    return loss, pred, opt.gflat.clone(), opt


@pytest.mark.parametrize("B,roll,p,bce", [(3, 0, 0.0, False), (5, 5, 0.3, False), (37, -7, 0.3, False), (300, 11, 0.3, False),
                                          (16, -3, 0.3, True), (149, 0, 0.5, False)])
def test_critic_bf16_step_vs_oracle(ops, B, roll, p, bce):
    csd, X, y, masks = _case(B, p, seed=B)
    loss_r, pred_r, grads_r = _oracle(csd, X, y, masks, roll, bce)
    loss_q, pred_q, grads_q = _oracle(csd, X, y, masks, roll, bce, q=torch_ref.quant_bf16)
    c = _critic(csd, p)
    yt = torch.from_numpy(y if not bce else (y > 0.5).astype(np.float32)).to(DEV)
    loss, pred, g, _ = _step(ops, c, torch.from_numpy(X).to(DEV), yt, roll, _dev_masks(masks), bce=bce)
    pred, g = pred.cpu().numpy(), g.cpu().numpy()
    names = [k for k, _ in c.named_parameters()]
    sizes = [v.numel() for v in c.parameters()]
    parts = dict(zip(names, np.split(g, np.cumsum(sizes)[:-1])))
    # (1) operand-precision oracle: same decisions; the remaining difference is the bf16 rounding of the gradients that the
    # kernel back-propagates through bf16 planes (each rounding 2^-9 relative, independent)
    assert np.abs(pred - pred_q).max() <= 1e-3, np.abs(pred - pred_q).max()
    assert abs(loss.item() - loss_q) <= 2e-3 * abs(loss_q) + 1e-6, (loss.item(), loss_q)
    gq = np.concatenate([grads_q[k].ravel() for k in names])
    errs_q = {k: _rel(parts[k], grads_q[k].ravel()) for k in names}
    tot_q = _rel(g, gq)
    assert tot_q <= 1e-2 and max(errs_q.values()) <= 3e-2, ("operand-precision oracle", tot_q, errs_q)
    # (2) the reference arithmetic
    assert np.abs(pred - pred_r).max() <= 1e-2, np.abs(pred - pred_r).max()
    assert abs(loss.item() - loss_r) <= 2e-2 * abs(loss_r) + 1e-6, (loss.item(), loss_r)
    gr = np.concatenate([grads_r[k].ravel() for k in names])
    errs = {k: _rel(parts[k], grads_r[k].ravel()) for k in names}
    tot = _rel(g, gr)
    # ReLU / arg-max flips under bf16 operand rounding move gradients norm-wise by sqrt(flip rate): the same two-level bounds as
    # the bf16 Hourglass step (tests/test_gpu_hg.py::test_hg_fused_step_vs_oracle)
    assert tot <= 1.5e-1, (tot, errs)
    if B >= 30:                  # with a handful of frames a single flipped decision is a large share of one small tensor's gradient
        assert max(errs.values()) <= 2.5e-1, errs


def test_critic_bf16_is_reproducible_and_scales_with_loss_grad(ops):
    B = 41
    csd, X, y, masks = _case(B, 0.3, seed=77)
    Xd, yd, dm = torch.from_numpy(X).to(DEV), torch.from_numpy(y).to(DEV), _dev_masks(masks)
    l1, p1, g1, _ = _step(ops, _critic(csd, 0.3), Xd, yd, 6, dm)
    l2, p2, g2, _ = _step(ops, _critic(csd, 0.3), Xd, yd, torch.tensor([6], dtype=torch.int32, device=DEV), dm)
    assert torch.equal(p1, p2) and torch.equal(g1, g2)             # fixed-order sums: bit-reproducible
    assert abs(l1.item() - l2.item()) <= 1e-6 * abs(l1.item())
    _, _, g3, _ = _step(ops, _critic(csd, 0.3), Xd, yd, 6, dm, loss_grad=0.5)
    # the scale enters before the bf16 roundings of the back-propagated planes: equal up to those roundings
    assert _rel(g3.cpu().numpy(), 0.5 * g1.cpu().numpy()) <= 5e-3


def test_critic_bf16_close_to_tf32_kernel(ops):
    """The two whole-step kernels on the same inputs: same loss / predictions / gradients up to operand precision."""
    from cgs_b200.train_handler import FlatAdam
    B = 150
    csd, X, y, masks = _case(B, 0.3, seed=9)
    Xd, yd, dm = torch.from_numpy(X).to(DEV), torch.from_numpy(y).to(DEV), _dev_masks(masks)
    lb, pb, gb, _ = _step(ops, _critic(csd, 0.3), Xd, yd, 3, dm)
    c = _critic(csd, 0.3)
    opt = FlatAdam(c.parameters())
    opt.zero_grad()
    lt, pt = ops.critic_train_fused(c, Xd, yd, 3, dm)
    opt.flush_partials()
    assert (pb - pt).abs().max().item() <= 1e-2
    assert abs(lb.item() - lt.item()) <= 2e-2 * abs(lt.item())
    assert _rel(gb.cpu().numpy(), opt.gflat.cpu().numpy()) <= 1.5e-1       # two sets of decision flips


@pytest.mark.parametrize("B", [5, 256, 300])
def test_critic_bf16_in_kernel_adam_matches_adam_kernel(ops, B):
    """Adam applied inside the bf16 kernel (grid barrier + per-CTA slice: critic_tail.cuh, shared with the TF32 kernel)
    == partial hand-over + Adam kernel."""
    from cgs_b200.train_handler import FlatAdam
    csd, X, y, masks = _case(B, 0.3, seed=40 + B)
    Xd, yd, dm = torch.from_numpy(X).to(DEV), torch.from_numpy(y).to(DEV), _dev_masks(masks)
    ca, cb = _critic(csd, 0.3), _critic(csd, 0.3)
    oa, ob = FlatAdam(ca.parameters()), FlatAdam(cb.parameters())
    la, lb = [], []
    for step in range(3):
        for c, o, fuse, ls in ((ca, oa, True, la), (cb, ob, False, lb)):
            o.zero_grad()
            l, _ = ops.critic_train_fused(c, Xd, yd, step, dm, fuse_adam=fuse, bf16=True)
            assert o.adam_done_in_kernel == fuse
            o.step()
            ls.append(l.item())
    assert oa.barrier_ok()
    assert int(oa.step_count[0]) == 3 and int(ob.step_count[0]) == 3
    assert np.allclose(la, lb, rtol=2e-3), (la, lb)
    # parameters differ by the fp32 summation order only, until a bf16 operand rounding of a weight lands on the other side
    assert (oa.flat - ob.flat).abs().max().item() <= 1e-4, (oa.flat - ob.flat).abs().max().item()
    assert (oa.flat - ob.flat).abs().mean().item() <= 1e-6
    assert float(oa.gflat.abs().max()) == 0.0


def test_critic_bf16_in_kernel_dropout_is_the_mask_kernels_stream(ops):
    B = 37
    csd, X, y, _ = _case(B, 0.3, seed=5)
    Xd, yd = torch.from_numpy(X).to(DEV), torch.from_numpy(y).to(DEV)
    state, seed = torch.zeros(2, dtype=torch.int64, device=DEV), 0x1234567
    masks = ops.dropout_masks([(B, 8, 8, 8), (B, 4, 4, 16), (B, 32)], 0.3, seed, state)
    assert int(state[0]) == 1
    _, p1, g1, _ = _step(ops, _critic(csd, 0.3), Xd, yd, 2, tuple(masks))
    state.zero_()
    _, p2, g2, _ = _step(ops, _critic(csd, 0.3), Xd, yd, 2, (None, None, None), rng=(0.3, seed, state))
    assert int(state[0]) == 1 and int(state[1]) == 0
    assert torch.equal(p1, p2) and torch.equal(g1, g2)
    _, p3, _, _ = _step(ops, _critic(csd, 0.3), Xd, yd, 2, (None, None, None), rng=(0.3, seed, state))
    assert int(state[0]) == 2 and not torch.equal(p2, p3)


def test_critic_bf16_batch_linearity_at_8192(ops):
    """Full-size property: the gradient of a batch is the frame-count-weighted mean of its halves' gradients (the kernel is
    per-frame independent); predictions do not depend on which CTA scored the frame."""
    B = 8192
    csd, X, y, _ = _case(512, 0.0, seed=3)
    X = np.tile(X, (16, 1, 1, 1)); y = np.tile(y, 16)
    X[1::2] = X[1::2, :, ::-1]
    Xd, yd = torch.from_numpy(np.ascontiguousarray(X)).to(DEV), torch.from_numpy(y).to(DEV)
    none = (None, None, None)
    l, p, g, _ = _step(ops, _critic(csd, 0.0), Xd, yd, 0, none)
    la, pa, ga, _ = _step(ops, _critic(csd, 0.0), Xd[:B // 2], yd[:B // 2], 0, none)
    lb, pb, gb, _ = _step(ops, _critic(csd, 0.0), Xd[B // 2:], yd[B // 2:], 0, none)
    assert torch.equal(p, torch.cat([pa, pb]))
    assert abs(l.item() - 0.5 * (la.item() + lb.item())) <= 1e-5 * abs(l.item())
    # gscale = loss_grad / B enters before the bf16 roundings: halves were back-propagated at twice the scale
    assert _rel(g.cpu().numpy(), 0.5 * (ga + gb).cpu().numpy()) <= 5e-3


def test_critic_bf16_handler_step_and_loss_curve(ops):
    """Handler.critic_pipe through the bf16 kernel from the reference's initial weights: the deterministic regime of the
    reference loss curve (first 40 steps, epoch-1 median) within bf16 noise; and the Handler really took the bf16 path."""
    from cgs_b200.train_handler import Handler, parse_args
    d = load_golden("loops_c1.npz")
    N = 6000
    X, Y, I = synth.synthetic_frames(N, seed=0)
    H = Handler(parse_args(["--dropout", "0", "--shift", "0", "--cepochs", "2", "--saveevery", "100", "--model", "/tmp/cgs_bf16_loop"]),
                device=DEV)
    assert H.critic_bf16
    H.args.cload = False
    H.critic.load_state_dict({k[len("init.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("init.c.")})
    H.critic.to(DEV)
    Xt, Yt = torch.from_numpy(X), torch.from_numpy(Y).t()
    H.train_loader = [(Xt[i:i + 64], Yt[i:i + 64], None) for i in range(0, N, 64)]
    H.critic_pipe()
    closs, ref = np.array(H.closs_log), d["closs"]
    early = np.abs(closs[:40] - ref[:40]) / ref[:40]
    assert early.max() < 0.03, early.max()
    assert abs(np.median(closs[:94]) - np.median(ref[:94])) <= 0.02 * np.median(ref[:94])


@pytest.mark.parametrize("bf16", [True, False])
def test_critic_tensor_core_loss_curve_inside_reference_envelope(ops, bf16, monkeypatch):
    """All 11 epochs of critic_pipe (1034 steps) through the whole-step kernels - bf16 (csrc/hg_critic.cu) and TF32
    (csrc/critic_fused.cu) - from the reference's initial weights: every per-epoch median loss inside the band the reference itself
    spans when re-run from 1e-6-perturbed weights (north_star: "loss curves within 1 % over 1k steps" is not satisfiable by the
    reference against itself beyond epoch 1; see helpers.assert_in_epoch_band), the first 40 steps pointwise, and convergence."""
    from helpers import assert_in_epoch_band
    from cgs_b200.train_handler import Handler, parse_args
    d = load_golden("loops_c1.npz")
    N = 6000
    X, Y, I = synth.synthetic_frames(N, seed=0)
    H = Handler(parse_args(["--dropout", "0", "--shift", "0", "--cepochs", "11", "--saveevery", "100", "--model", "/tmp/cgs_env_loop"]), device=DEV)
    H.critic_bf16 = bf16
    H.args.cload = False
    H.critic.load_state_dict({k[len("init.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("init.c.")})
    H.critic.to(DEV)
    Xt, Yt = torch.from_numpy(X), torch.from_numpy(Y).t()
    H.train_loader = [(Xt[i:i + 64], Yt[i:i + 64], None) for i in range(0, N, 64)]
    H.critic_pipe()
    closs, ref = np.array(H.closs_log), d["closs"]
    early = np.abs(closs[:40] - ref[:40]) / ref[:40]
    assert early.max() < 0.03, early.max()
    ours, lo, hi = assert_in_epoch_band(closs, slack=0.03, what="bf16" if bf16 else "tf32", operand_precision=True)
    assert ours[-1] < 0.02 * ours[0]
