"""GPU: the tcgen05 / TMEM TF32 convolution path against a torch fp32 CPU reference.
Tolerance: TF32 operands (10-bit mantissa), fp32 accumulation -> 2e-3 of the tensor scale; the
model-level bound BASELINE.json states for this path is mask abs error <= 2e-2 and IoU >= 0.99."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def close(ours, ref, what, tol=2e-3):
    ours = ours.detach().cpu().double().numpy()
    ref = ref.detach().cpu().double().numpy()
    assert ours.shape == ref.shape, f"{what}: {ours.shape} vs {ref.shape}"
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(ours - ref).max()
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.2e})"


def close_l2(ours, ref, what, tol):
    """Norm-wise comparison for gradients routed through max-pool arg-max decisions: under TF32 rounding a
    near-tie can pick the other window position, which moves one gradient entry by its full value."""
    ours = ours.detach().cpu().double().numpy()
    ref = ref.detach().cpu().double().numpy()
    rel = np.linalg.norm(ours - ref) / max(np.linalg.norm(ref), 1e-30)
    assert rel <= tol, f"{what}: relative L2 error {rel:.3e} > {tol}"


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


@pytest.fixture()
def ops():
    import cgs_b200.ops as o
    from cgs_b200 import _lib
    o.set_precision("tf32")
    yield o
    o.set_precision("fp32")
    torch.cuda.synchronize()
    assert _lib.lib().cgs_tc_status() == 0, "a tcgen05 kernel timed out on its completion barrier"


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 64, 3, 8), (3, 32, 8, 8), (5, 16, 8, 8), (1, 64, 3, 40), (2, 32, 40, 40),
                                          (2, 16, 40, 80), (37, 16, 8, 16)])
def test_tc_encblock(ops, B, H, Cin, Cout):
    x = rnd(B, Cin, H, H, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=(3.0 / (Cin * 9)) ** 0.5)
    b = rnd(Cout, seed=3, scale=0.1)
    de = rnd(B, Cout, H // 2, H // 2, seed=5)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    er = F.max_pool2d(F.relu(F.conv2d(xr, wr, br, padding=1)), 2)
    er.backward(de)
    xo = nhwc(x).cuda().requires_grad_()
    wo, bo = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    eo = ops.EncBlock.apply(xo, None, wo, bo)
    eo.backward(nhwc(de).cuda())
    close(nchw(eo), er, "e")
    # arg-max flips under TF32 rounding move single gradient entries: compare in aggregate
    close_l2(nchw(xo.grad), xr.grad, "dx", 1e-1)
    close_l2(wo.grad, wr.grad, "dw", 6e-2)


@pytest.mark.parametrize("B,H,C0,C1,Cout,leaky", [(2, 16, 8, 8, 8, False), (2, 32, 8, 8, 8, False), (2, 64, 3, 8, 16, True),
                                                  (2, 64, 3, 40, 16, True), (3, 32, 40, 40, 40, False)])
def test_tc_decblock(ops, B, H, C0, C1, Cout, leaky):
    skip, up = rnd(B, C0, H, H, seed=1), rnd(B, C1, H // 2, H // 2, seed=2)
    Cin = C0 + C1
    w = rnd(Cout, Cin, 3, 3, seed=3, scale=(3.0 / (Cin * 9)) ** 0.5)
    b = rnd(Cout, seed=4, scale=0.1)
    dout = rnd(B, Cout, H, H, seed=5)
    sr, ur, wr, br = (t.clone().requires_grad_() for t in (skip, up, w, b))
    o = F.conv2d(torch.cat((sr, F.interpolate(ur, scale_factor=2, mode="nearest")), 1), wr, br, padding=1)
    if leaky:
        o = F.leaky_relu(o, 0.01)
    o.backward(dout)
    so, uo = nhwc(skip).cuda().requires_grad_(), nhwc(up).cuda().requires_grad_()
    wo, bo = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    oo = ops.DecBlock.apply(so, uo, wo, bo, 1, leaky)
    oo.backward(nhwc(dout).cuda())
    close(nchw(oo), o, "out")
    if leaky:    # LeakyReLU' flips between 1 and 0.01 where TF32 rounding changes the sign of a near-zero output
        close_l2(nchw(so.grad), sr.grad, "dskip", 1e-1)
        close_l2(nchw(uo.grad), ur.grad, "dup", 1e-1)
        close_l2(wo.grad, wr.grad, "dw", 6e-2)
    else:
        close(nchw(so.grad), sr.grad, "dskip", tol=4e-3)
        close(nchw(uo.grad), ur.grad, "dup", tol=4e-3)
        close(wo.grad, wr.grad, "dw", tol=4e-3)


def test_tc_maskhead(ops):
    m = rnd(2, 16, 64, 64, seed=1)
    w, b = rnd(1, 16, 3, 3, seed=2, scale=0.3), rnd(1, seed=3, scale=0.1)
    dz = rnd(2, 1, 64, 64, seed=4)
    mr, wr, br = (t.clone().requires_grad_() for t in (m, w, b))
    z = torch.sigmoid(F.conv2d(mr, wr, br, padding=1))
    z.backward(dz)
    mo = nhwc(m).cuda().requires_grad_()
    wo, bo = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    zo, hard = ops.MaskHead.apply(mo, wo, bo, 0.5)
    zo.backward(nhwc(dz).cuda())
    close(nchw(zo), z, "z")
    assert torch.equal(hard.bool(), zo >= 0.5)
    close(nchw(mo.grad), mr.grad, "dm", tol=4e-3)


def test_tc_model_level_mask_bound():
    """Whole inference path in TF32 on reference-trained weights: |mask - reference| <= 2e-2, IoU >= 0.99 @0.1."""
    from cgs_b200 import ops as o
    from cgs_b200.train_handler import Handler, parse_args
    import cgs_b200.synth as synth
    from helpers import load_golden
    d = load_golden("loops_c1.npz")
    H = Handler(parse_args(["--binarymaskthreshold", "0.1"]), device="cuda")
    H.critic.load_state_dict({k[len("trained.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.c.")})
    H.masker.load_state_dict({k[len("trained.m."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.m.")})
    X, _, _ = synth.synthetic_frames(6000, seed=0)
    o.set_precision("tf32")
    try:
        preds, M, hard = H.segment_arrays(X[:32])
    finally:
        o.set_precision("fp32")
    assert np.abs(M - d["proc_mask"]).max() <= 2e-2, np.abs(M - d["proc_mask"]).max()
    ref_hard = np.unpackbits(d["proc_hard"])[:hard.size].reshape(hard.shape).astype(bool)
    inter, union = (hard & ref_hard).sum(), (hard | ref_hard).sum()
    assert inter / union >= 0.99, inter / union
    assert np.abs(preds - d["proc_pred"].reshape(-1)).max() <= 2e-2
