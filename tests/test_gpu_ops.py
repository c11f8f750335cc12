"""GPU: every C-ABI kernel against a plain torch fp32 CPU reference of the same op
(rtol 1e-4 — the fp32-path tolerance BASELINE.json states — plus a scale-aware atol)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def close(ours, ref, what, rtol=RTOL, arel=2e-5):
    ours = ours.detach().cpu().double().numpy()
    ref = ref.detach().cpu().double().numpy()
    assert ours.shape == ref.shape, f"{what}: {ours.shape} vs {ref.shape}"
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(ours - ref)
    bad = err > rtol * np.abs(ref) + arel * scale
    assert not bad.any(), f"{what}: {bad.sum()}/{bad.size} bad, max err {err.max():.3e} (scale {scale:.3e})"


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2)


@pytest.fixture(scope="module")
def ops():
    import cgs_b200.ops as o
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


@pytest.mark.parametrize("B,H,Cin,Cout,use_mask", [
    (3, 64, 3, 8, False), (2, 32, 8, 8, False), (5, 16, 8, 8, False), (7, 8, 8, 16, True),
    (1, 64, 3, 40, False), (2, 32, 40, 40, False), (3, 8, 40, 80, True), (33, 8, 16, 24, True), (2, 16, 5, 12, False)])
def test_encblock_fwd_bwd(ops, B, H, Cin, Cout, use_mask):
    x = rnd(B, Cin, H, H, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, scale=(3.0 / (Cin * 9)) ** 0.5)
    b = rnd(Cout, seed=3, scale=0.1)
    mask = (torch.rand(B, Cin, H, H, generator=torch.Generator().manual_seed(4)) > 0.3).float() / 0.7 if use_mask else None
    de = rnd(B, Cout, H // 2, H // 2, seed=5)
    # reference
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    xin = xr * mask if use_mask else xr
    er = F.max_pool2d(F.relu(F.conv2d(xin, wr, br, padding=1)), 2)
    er.backward(de)
    # ours
    dev = "cuda"
    xo = nhwc(x).to(dev).requires_grad_()
    wo, bo = w.to(dev).requires_grad_(), b.to(dev).requires_grad_()
    mo = nhwc(mask).to(dev) if use_mask else None
    eo = ops.EncBlock.apply(xo, mo, wo, bo)
    eo.backward(nhwc(de).to(dev))
    close(nchw(eo), er, "e")
    close(wo.grad, wr.grad, "dw")
    close(bo.grad, br.grad, "db")
    close(nchw(xo.grad), xr.grad, "dx")


def test_pool_tie_first_max_wins(ops):
    # all-equal positive windows: ATen routes the gradient to window position 0 (SURVEY.md §7 hard part 5)
    B, H, C = 2, 8, 8
    x = torch.zeros(B, 3, H, H)
    w = torch.zeros(C, 3, 3, 3)
    b = torch.full((C,), 0.5)
    xr, br = x.clone().requires_grad_(), b.clone().requires_grad_()
    wr = w.clone().requires_grad_()
    er = F.max_pool2d(F.relu(F.conv2d(xr, wr, br, padding=1)), 2)
    de = rnd(B, C, H // 2, H // 2, seed=9)
    er.backward(de)
    xo = nhwc(x).cuda().requires_grad_()
    wo, bo = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    eo = ops.EncBlock.apply(xo, None, wo, bo)
    eo.backward(nhwc(de).cuda())
    close(nchw(eo), er, "e")
    close(bo.grad, br.grad, "db")
    close(wo.grad, wr.grad, "dw")
    close(nchw(xo.grad), xr.grad, "dx")


@pytest.mark.parametrize("B,H,C0,C1,Cout,shift,leaky", [
    (3, 4, 16, 32, 16, 2, False), (2, 8, 8, 16, 8, 1, False), (2, 16, 8, 8, 8, 1, False), (2, 32, 8, 8, 8, 1, False),
    (2, 64, 3, 8, 16, 1, True), (1, 4, 80, 160, 80, 2, False), (2, 64, 3, 40, 16, 1, True), (35, 8, 40, 80, 40, 1, False)])
def test_decblock_fwd_bwd(ops, B, H, C0, C1, Cout, shift, leaky):
    skip = rnd(B, C0, H, H, seed=1)
    up = rnd(B, C1, H >> shift, H >> shift, seed=2)
    Cin = C0 + C1
    w = rnd(Cout, Cin, 3, 3, seed=3, scale=(3.0 / (Cin * 9)) ** 0.5)
    b = rnd(Cout, seed=4, scale=0.1)
    dout = rnd(B, Cout, H, H, seed=5)
    sr, ur, wr, br = (t.clone().requires_grad_() for t in (skip, up, w, b))
    u = ur
    for _ in range(shift):
        u = F.interpolate(u, scale_factor=2, mode="nearest")
    o = F.conv2d(torch.cat((sr, u), 1), wr, br, padding=1)
    if leaky:
        o = F.leaky_relu(o, 0.01)
    o.backward(dout)
    so, uo = nhwc(skip).cuda().requires_grad_(), nhwc(up).cuda().requires_grad_()
    wo, bo = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    oo = ops.DecBlock.apply(so, uo, wo, bo, shift, leaky)
    oo.backward(nhwc(dout).cuda())
    close(nchw(oo), o, "out")
    close(wo.grad, wr.grad, "dw")
    close(bo.grad, br.grad, "db")
    close(nchw(so.grad), sr.grad, "dskip")
    close(nchw(uo.grad), ur.grad, "dup")


@pytest.mark.parametrize("B,Cm", [(2, 16), (3, 8)])
def test_maskhead_fwd_bwd(ops, B, Cm):
    m = rnd(B, Cm, 64, 64, seed=1)
    w = rnd(1, Cm, 3, 3, seed=2, scale=0.3)
    b = rnd(1, seed=3, scale=0.1)
    dz = rnd(B, 1, 64, 64, seed=4)
    mr, wr, br = (t.clone().requires_grad_() for t in (m, w, b))
    z = torch.sigmoid(F.conv2d(mr, wr, br, padding=1))
    z.backward(dz)
    mo = nhwc(m).cuda().requires_grad_()
    wo, bo = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    zo, hard = ops.MaskHead.apply(mo, wo, bo, 0.5)
    zo.backward(nhwc(dz).cuda())
    close(nchw(zo), z, "z")
    # the threshold must agree with the kernel's own z everywhere, and with the reference away from the edge
    assert torch.equal(hard.bool(), zo >= 0.5)
    far = (z - 0.5).abs() > 1e-5
    assert torch.equal(nchw(hard).cpu().bool()[far], (z >= 0.5)[far])
    close(wo.grad, wr.grad, "dw")
    close(bo.grad, br.grad, "db")
    close(nchw(mo.grad), mr.grad, "dm")


@pytest.mark.parametrize("B,C3,NB,use_masks,use_de4", [(5, 16, 32, True, True), (37, 16, 32, False, False),
                                                       (3, 80, 160, True, True), (16, 32, 64, True, False)])
def test_head_fwd_bwd(ops, B, C3, NB, use_masks, use_de4):
    e3 = rnd(B, C3, 4, 4, seed=1).abs()
    w14 = rnd(NB, C3, 4, 4, seed=2, scale=(3.0 / (C3 * 16)) ** 0.5)
    b14 = rnd(NB, seed=3, scale=0.1)
    w1 = rnd(NB, NB, seed=4, scale=(3.0 / NB) ** 0.5)
    b1 = rnd(NB, seed=5, scale=0.1)
    w2 = rnd(1, NB, seed=6, scale=(3.0 / NB) ** 0.5)
    b2 = rnd(1, seed=7, scale=0.1)
    g = torch.Generator().manual_seed(8)
    m3 = (torch.rand(B, C3, 4, 4, generator=g) > 0.3).float() / 0.7 if use_masks else None
    mv = (torch.rand(B, NB, generator=g) > 0.3).float() / 0.7 if use_masks else None
    dpred = rnd(B, 1, seed=9)
    de4 = rnd(B, NB, 1, 1, seed=10) if use_de4 else None
    ps = [t.clone().requires_grad_() for t in (e3, w14, b14, w1, b1, w2, b2)]
    x = ps[0] * m3 if use_masks else ps[0]
    h = F.relu(F.conv2d(x, ps[1], ps[2]))
    v = F.relu(F.linear(h.flatten(1), ps[3], ps[4]))
    vd = v * mv if use_masks else v
    pred = torch.sigmoid(F.linear(vd, ps[5], ps[6]))
    torch.autograd.backward([pred, h] if use_de4 else [pred], [dpred, de4] if use_de4 else [dpred])
    po = [nhwc(e3).cuda().requires_grad_()] + [t.cuda().requires_grad_() for t in (w14, b14, w1, b1, w2, b2)]
    m3o = nhwc(m3).cuda() if use_masks else None
    mvo = mv.cuda() if use_masks else None
    predo, e4o = ops.Head.apply(po[0], m3o, mvo, *po[1:])
    if use_de4:
        torch.autograd.backward([predo, e4o], [dpred.cuda(), nhwc(de4).cuda()])
    else:
        predo.backward(dpred.cuda())
    close(predo, pred, "pred")
    close(nchw(e4o), h, "e4")
    names = ["de3", "dw14", "db14", "dw1", "db1", "dw2", "db2"]
    close(nchw(po[0].grad), ps[0].grad, "de3")
    for n, a, r in zip(names[1:], po[1:], ps[1:]):
        close(a.grad, r.grad, n)


@pytest.mark.parametrize("B,C2,C3,NB,use_masks,extra", [(5, 8, 16, 32, True, True), (37, 8, 16, 32, False, False),
                                                         (3, 16, 32, 64, True, True), (2, 24, 48, 96, True, False)])
def test_fused_tail_fwd_bwd(ops, B, C2, C3, NB, use_masks, extra):
    """features[9..15] + crit in one kernel each way vs the same chain in torch (fp32)."""
    assert ops.tail_supported(B, C2, C3, NB)
    e2 = rnd(B, C2, 8, 8, seed=1).abs()
    w3 = rnd(C3, C2, 3, 3, seed=2, scale=(3.0 / (C2 * 9)) ** 0.5)
    b3 = rnd(C3, seed=3, scale=0.1)
    w14 = rnd(NB, C3, 4, 4, seed=4, scale=(3.0 / (C3 * 16)) ** 0.5)
    b14 = rnd(NB, seed=5, scale=0.1)
    w1 = rnd(NB, NB, seed=6, scale=(3.0 / NB) ** 0.5)
    b1 = rnd(NB, seed=7, scale=0.1)
    w2 = rnd(1, NB, seed=8, scale=(3.0 / NB) ** 0.5)
    b2 = rnd(1, seed=9, scale=0.1)
    g = torch.Generator().manual_seed(10)
    mk = lambda *s: (torch.rand(*s, generator=g) > 0.3).float() / 0.7
    m2, m3, mv = (mk(B, C2, 8, 8), mk(B, C3, 4, 4), mk(B, NB)) if use_masks else (None, None, None)
    dpred, de3, de4 = rnd(B, 1, seed=11), rnd(B, C3, 4, 4, seed=12), rnd(B, NB, 1, 1, seed=13)
    ps = [t.clone().requires_grad_() for t in (e2, w3, b3, w14, b14, w1, b1, w2, b2)]
    x = ps[0] * m2 if use_masks else ps[0]
    e3 = F.max_pool2d(F.relu(F.conv2d(x, ps[1], ps[2], padding=1)), 2)
    x3 = e3 * m3 if use_masks else e3
    h = F.relu(F.conv2d(x3, ps[3], ps[4]))
    v = F.relu(F.linear(h.flatten(1), ps[5], ps[6]))
    pred = torch.sigmoid(F.linear(v * mv if use_masks else v, ps[7], ps[8]))
    if extra:
        torch.autograd.backward([pred, e3, h], [dpred, de3, de4])
    else:
        pred.backward(dpred)
    po = [nhwc(e2).cuda().requires_grad_()] + [t.cuda().requires_grad_() for t in (w3, b3, w14, b14, w1, b1, w2, b2)]
    dm = lambda t: None if t is None else (nhwc(t).cuda() if t.dim() == 4 else t.cuda())
    predo, e3o, e4o = ops.Tail.apply(po[0], dm(m2), dm(m3), dm(mv), *po[1:])
    if extra:
        torch.autograd.backward([predo, e3o, e4o], [dpred.cuda(), nhwc(de3).cuda(), nhwc(de4).cuda()])
    else:
        predo.backward(dpred.cuda())
    close(predo, pred, "pred")
    close(nchw(e3o), e3, "e3")
    close(nchw(e4o), h, "e4")
    close(nchw(po[0].grad), ps[0].grad, "de2")
    for n, a, r in zip(["dw3", "db3", "dw14", "db14", "dw1", "db1", "dw2", "db2"], po[1:], ps[1:]):
        close(a.grad, r.grad, n)


@pytest.mark.parametrize("B,K,N", [(5, 32, 32), (40, 160, 160), (1, 7, 3)])
def test_dense_fwd_bwd(ops, B, K, N):
    x, w, b, do = rnd(B, K, seed=1), rnd(N, K, 1, 1, seed=2, scale=0.2), rnd(N, seed=3), rnd(B, N, seed=4)
    xr, wr, br = (t.clone().requires_grad_() for t in (x, w, b))
    o = F.linear(xr, wr.view(N, K), br)
    o.backward(do)
    xo = x.view(B, 1, 1, K).cuda().requires_grad_()
    wo, bo = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    oo = ops.Dense.apply(xo, wo, bo)
    oo.backward(do.view(B, 1, 1, N).cuda())
    close(oo.view(B, N), o, "out")
    close(xo.grad.view(B, K), xr.grad, "dx")
    close(wo.grad, wr.grad, "dw")
    close(bo.grad, br.grad, "db")


def test_occlude_and_losses(ops):
    B = 3
    A, Bf = rnd(B, 64, 64, 3, seed=1).abs(), rnd(B, 64, 64, 3, seed=2).abs()
    Z = rnd(B, 64, 64, 1, seed=3).abs()
    g = rnd(B, 64, 64, 3, seed=4)
    Ar, Br, Zr = (t.clone().requires_grad_() for t in (A, Bf, Z))
    out = Ar * (1 - Zr) + Zr * Br
    out.backward(g)
    Ao, Bo, Zo = (t.cuda().requires_grad_() for t in (A, Bf, Z))
    oo = ops.occlude(Ao, Bo, Zo)
    oo.backward(g.cuda())
    close(oo, out, "occlude")
    close(Zo.grad, Zr.grad, "dz")
    close(Ao.grad, Ar.grad, "da")
    close(Bo.grad, Br.grad, "db")
    # prediction losses
    p, t = rnd(37, seed=5).abs() * 0.98 + 0.01, rnd(37, seed=6).abs()
    for bce in (False, True):
        pr = p.clone().requires_grad_()
        lr = F.binary_cross_entropy(pr, t) if bce else F.mse_loss(pr, t)
        (3 * lr).backward()
        po = p.cuda().requires_grad_()
        lo = ops.pred_loss(po, t.cuda(), bce=bce)
        (3 * lo).backward()
        close(lo, lr, f"loss bce={bce}")
        close(po.grad, pr.grad, f"dloss bce={bce}")
    # mask regulariser, static and per-frame valuefak
    vp = rnd(B, seed=7).abs()
    for static in (True, False):
        for l1, l2 in ((0.5, 0.0), (0.0, 0.7)):
            zr = Z.clone().requires_grad_()
            vf = 1 if static else 1 - vp.view(-1, 1, 1, 1)
            lr = l1 * F.l1_loss(vf * zr, torch.zeros_like(zr)) + l2 * F.mse_loss(vf * zr, torch.zeros_like(zr))
            lr.backward()
            zo = Z.cuda().requires_grad_()
            lo = ops.mask_reg(zo, None if static else vp.cuda(), l1=l1, l2=l2)
            lo.backward()
            close(lo, lr, "reg")
            close(zo.grad, zr.grad, "dreg")


def test_frames_to_float_threshold_adam(ops):
    from oracle import torch_ref
    X = torch.randint(0, 256, (5, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    for xs, left in ((0, True), (5, True), (11, False), (0, False)):
        ref = torch_ref.shift_batch(X, xs, left).float() / 255.0
        got = ops.frames_to_float(X.cuda(), xs if left else -xs)
        assert torch.equal(got.cpu(), ref)      # bit-exact: integer gather + one fp32 division
    z = torch.rand(1000, generator=torch.Generator().manual_seed(1))
    z[::7] = 0.1
    assert torch.equal(ops.threshold(z.cuda(), 0.1).cpu().bool(), z >= 0.1)
    assert torch.equal(ops.threshold(z.cuda(), 0.1, strict=True).cpu().bool(), z > 0.1)
    # Adam vs torch.optim.Adam defaults
    p0 = rnd(1000, seed=2)
    q = p0.clone().requires_grad_()
    opt = torch.optim.Adam([q])
    p, m, v = p0.cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    step = torch.zeros(2, dtype=torch.int32).cuda()
    for it in range(6):
        g = rnd(1000, seed=10 + it)
        q.grad = g.clone()
        opt.step()
        gd = g.cuda()
        ops.adam_step(p, gd, m, v, step, clear_grad=(it % 2 == 0))
        assert (gd.abs().sum().item() == 0) == (it % 2 == 0)
    assert step.cpu().tolist() == [6, 0]
    close(p, q, "adam", rtol=1e-6, arel=1e-7)


def test_dropout_masks_kernel(ops):
    state = torch.zeros(2, dtype=torch.int64, device="cuda")
    p = 0.3
    shapes = [(64, 8, 8, 8), (64, 4, 4, 16), (64, 32)]
    a = ops.dropout_masks(shapes, p, 1234, state)
    b = ops.dropout_masks(shapes, p, 1234, state)
    assert int(state[0]) == 2 and int(state[1]) == 0
    for m, s in zip(a, shapes):
        assert tuple(m.shape) == s
        vals = torch.unique(m).cpu().tolist()
        assert all(abs(v) < 1e-12 or abs(v - 1 / (1 - p)) < 1e-6 for v in vals)
    keep = torch.cat([(m > 0).float().flatten() for m in a])
    n = keep.numel()
    assert abs(keep.mean().item() - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5
    assert not torch.equal(a[0], b[0])                      # the call counter advanced on the device
    state2 = torch.zeros(2, dtype=torch.int64, device="cuda")
    c = ops.dropout_masks(shapes, p, 1234, state2)
    assert all(torch.equal(x, y) for x, y in zip(a, c))     # same (seed, counter) -> same masks
    # graph replays draw fresh masks
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.dropout_masks(shapes, p, 7, state)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        out = ops.dropout_masks(shapes, p, 7, state)
    g.replay(); r1 = out[0].clone(); g.replay(); r2 = out[0].clone()
    assert not torch.equal(r1, r2)


def test_errors_are_loud(ops):
    from cgs_b200._lib import CgsError
    x = torch.zeros(1, 6, 6, 3).cuda()      # H not a power of two
    w, b = torch.zeros(8, 3, 3, 3).cuda(), torch.zeros(8).cuda()
    with pytest.raises(CgsError):
        ops.EncBlock.apply(x, None, w, b)
    with pytest.raises(CgsError):
        ops.EncBlock.apply(torch.zeros(1, 8, 8, 3), None, w.cpu(), b.cpu())   # CPU tensors: no fallback


def test_iou_counts_exact():
    """cgs_iou_counts == np.sum(A & B), np.sum(A | B) of the reference's get_iou (main.py:1265-1270), exactly, for both
    threshold conventions, ragged sizes and accumulation over several calls."""
    from cgs_b200 import ops
    rng = np.random.default_rng(0)
    counts = torch.zeros(2, dtype=torch.int64, device="cuda")
    ti = tu = 0
    for n, strict in ((1, True), (1000, True), (64 * 64 * 37 + 5, False), (4096 * 300, True)):
        z = rng.random(n).astype(np.float32)
        z[::7] = 0.05                                   # values exactly at the threshold: > and >= must differ
        gt = (rng.random(n) < 0.3)
        hard = z > np.float32(0.05) if strict else z >= np.float32(0.05)
        ti += int((hard & gt).sum()); tu += int((hard | gt).sum())
        ops.iou_counts(torch.from_numpy(z).cuda(), torch.from_numpy(gt.astype(np.uint8)).cuda(), 0.05, counts, strict=strict)
    assert [int(v) for v in counts.cpu()] == [ti, tu]


def test_handler_eval_iou_matches_reference_formula():
    """Handler.eval_iou on the reference-trained checkpoint == get_iou(M > eval_thresh, GT) computed on the host from the
    same masks; synthetic ground truth = the painted trunk rectangle."""
    from cgs_b200.train_handler import Handler, parse_args
    import cgs_b200.synth as synth
    from helpers import load_golden
    d = load_golden("loops_c1.npz")
    H = Handler(parse_args(["--binarymaskthreshold", "0.1"]), device="cuda")
    H.critic.load_state_dict({k[len("trained.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.c.")})
    H.masker.load_state_dict({k[len("trained.m."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.m.")})
    X, _, _ = synth.synthetic_frames(300, seed=3)
    GT = (X[..., 0].astype(np.int32) >= 200)           # trunk pixels are painted (200..229, 140.., 60..), background < 120
    iou, inter, union = H.eval_iou(X, GT, batchsize=128)
    _, M, _ = H.segment_arrays(X)
    hard = M[:, 0] > np.float32(H.args.eval_thresh)
    assert (inter, union) == (int((hard & GT).sum()), int((hard | GT).sum()))
    assert iou == round(inter / union, 3) and 0.0 <= iou <= 1.0
