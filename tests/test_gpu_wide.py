"""GPU: the TMA-fed tcgen05 convolution kernels of the wide (chfak > 1) path (csrc/wide_tc.cu) against torch on the CPU.

Operands are bf16 (the kernel's storage format), accumulation fp32: the reference is evaluated in fp32 on the SAME bf16-rounded
operands (nets.py:170-183's Conv2d / ReLU / MaxPool2d / Dropout and their autograd backward), so what is left is summation order
and the bf16 rounding of the stored result (2^-9 relative)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(t):
    return t.to(torch.bfloat16).float()


def _rand(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g) * scale


SHAPES = [(3, 40, 40, 32, 32), (5, 40, 40, 16, 16), (4, 40, 80, 8, 8), (2, 16, 16, 32, 32), (2, 80, 48, 16, 8), (300, 40, 40, 16, 16)]


@pytest.mark.parametrize("B,Cin,Cout,H,W", SHAPES)
def test_wide_conv_plain_and_pool(B, Cin, Cout, H, W):
    from cgs_b200 import wide
    x, w, b = _bf(_rand(B, Cin, H, W, seed=1)), _rand(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5), _rand(Cout, seed=3, scale=0.1)
    ref = F.conv2d(x, _bf(w), b, padding=1)
    xp = wide.to_planar(x).to(DEV)
    out = wide.from_planar(wide.conv3x3(xp, w.to(DEV), b.to(DEV)).cpu())
    assert wide.status_ok()
    err = (out - ref).abs().max().item()
    assert err <= 6e-3 * ref.abs().max().item() + 1e-6, err
    (pk,) = wide.pack_weights([(w.to(DEV), False)])                  # operand tiles packed once by cgs_wide_pack: same bits
    assert torch.equal(wide.from_planar(wide.conv3x3(xp, None if False else w.to(DEV), b.to(DEV), packed=pk).cpu()), out)
    # ReLU + pool + arg-max (+ dropout mask)
    g = torch.Generator().manual_seed(4)
    mask = ((torch.rand(B, H // 2, W // 2, Cout, generator=g) >= 0.3).float() / 0.7)
    o, idx, f32 = wide.conv3x3(xp, w.to(DEV), b.to(DEV), epi=wide.EPI_RELU_POOL, mask=mask.to(DEV), want_f32=True)
    assert wide.status_ok()
    pr, pi = F.max_pool2d(F.relu(ref), 2, return_indices=True)
    o_ref = pr * mask.permute(0, 3, 1, 2)
    o, idx = wide.from_planar(o.cpu()), wide.from_planar(idx.cpu()).long()
    assert (o - o_ref).abs().max().item() <= 6e-3 * o_ref.abs().max().item() + 1e-6
    assert (f32.cpu() - o_ref).abs().max().item() <= 1e-4 * o_ref.abs().max().item() + 1e-6
    # arg-max: torch's flat index inside the H x W plane -> window position; compare where the window's maximum is unambiguous
    yy, xx = pi // W, pi % W
    pos = (yy % 2) * 2 + (xx % 2)
    r = F.relu(ref).reshape(B, Cout, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, Cout, H // 2, W // 2, 4)
    top2 = r.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 1e-4 * ref.abs().max()
    dead = pr <= 0
    assert (idx[dead] == 4).all()
    assert (idx[clear & ~dead] == pos[clear & ~dead]).all()
    assert ((idx[~dead] >= 0) & (idx[~dead] <= 3)).all()


@pytest.mark.parametrize("B,Cx,Cout,H,W", [(3, 40, 40, 16, 16), (4, 80, 40, 8, 8), (2, 40, 40, 32, 32), (2, 16, 16, 8, 8)])
def test_wide_dgrad_and_unpool(B, Cx, Cout, H, W):
    """Input gradient of a layer with weight [Cx, Cout, 3, 3] (x = its output gradient), then the pool / ReLU / dropout backward
    scatter into the full-resolution gradient of the layer below."""
    from cgs_b200 import wide
    dy, w = _bf(_rand(B, Cx, H, W, seed=5)), _rand(Cx, Cout, 3, 3, seed=6, scale=(9 * Cx) ** -0.5)
    ref = F.conv_transpose2d(dy, _bf(w), padding=1)                     # d/d input of conv2d(input, w, padding=1)
    dyp = wide.to_planar(dy).to(DEV)
    out = wide.from_planar(wide.conv3x3(dyp, w.to(DEV), transposed=True).cpu())
    assert wide.status_ok()
    assert (out - ref).abs().max().item() <= 6e-3 * ref.abs().max().item() + 1e-6
    g = torch.Generator().manual_seed(7)
    idx = torch.randint(0, 5, (B, Cout, H, W), generator=g)
    mask = ((torch.rand(B, H, W, Cout, generator=g) >= 0.3).float() / 0.7)
    idx_p = idx.reshape(B, Cout // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.uint8).to(DEV)
    up = wide.from_planar(wide.conv3x3(dyp, w.to(DEV), epi=wide.EPI_UNPOOL, transposed=True, idx_in=idx_p, mask=mask.to(DEV)).cpu())
    assert wide.status_ok()
    val = ref * mask.permute(0, 3, 1, 2)
    exp = torch.zeros(B, Cout, 2 * H, 2 * W)
    for pos in range(4):
        exp[:, :, pos // 2::2, pos % 2::2] = torch.where(idx == pos, val, torch.zeros(()))
    assert (up - exp).abs().max().item() <= 6e-3 * exp.abs().max().item() + 1e-6
    assert ((up != 0) <= (exp != 0)).all()


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(3, 40, 40, 32, 32), (5, 40, 40, 16, 16), (4, 40, 80, 8, 8), (2, 16, 16, 16, 16), (600, 40, 40, 16, 16)])
def test_wide_wgrad(B, Cin, Cout, H, W):
    from cgs_b200 import wide
    x, dy = _bf(_rand(B, Cin, H, W, seed=8)), _bf(_rand(B, Cout, H, W, seed=9, scale=1.0 / B))
    dw_ref = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, 3, 3), dy.double(), padding=1).float()
    db_ref = dy.double().sum((0, 2, 3)).float()
    dw = torch.ones(Cout, Cin, 3, 3, device=DEV)            # += semantics
    db = torch.full((Cout,), 2.0, device=DEV)
    wide.wgrad3x3(wide.to_planar(x).to(DEV), wide.to_planar(dy).to(DEV), dw, db)
    torch.cuda.synchronize()
    assert wide.status_ok()
    ew = (dw.cpu() - 1.0 - dw_ref).abs().max().item()
    eb = (db.cpu() - 2.0 - db_ref).abs().max().item()
    assert ew <= 2e-5 * dw_ref.abs().max().item() + 2e-6, (ew, dw_ref.abs().max().item())
    assert eb <= 2e-5 * db_ref.abs().max().item() + 2e-6, (eb, db_ref.abs().max().item())
    # bit-reproducible: fixed-order partial sums
    dw2 = torch.ones_like(dw); db2 = torch.full_like(db, 2.0)
    wide.wgrad3x3(wide.to_planar(x).to(DEV), wide.to_planar(dy).to(DEV), dw2, db2)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


# ---------------------------------------------------------------------------------------------------------------------
# features.0 kernels, the head GEMM, and the whole wide critic step
@pytest.mark.parametrize("B,C0,roll", [(3, 40, 0), (5, 16, 7), (150, 40, -3)])
def test_wide_conv0_fwd_and_wgrad(B, C0, roll):
    from cgs_b200 import wide
    import cgs_b200.synth as synth
    X, _, _ = synth.synthetic_frames(B, seed=B)
    Xr = np.roll(X, -roll, axis=2)
    x = _bf(torch.from_numpy(Xr).permute(0, 3, 1, 2).float() / 255.0)
    w, b = _rand(C0, 3, 3, 3, seed=2, scale=27 ** -0.5), _rand(C0, seed=3, scale=0.1)
    ref = F.conv2d(x, _bf(w), b, padding=1)
    pr, pi = F.max_pool2d(F.relu(ref), 2, return_indices=True)
    Xd = torch.from_numpy(X).to(DEV)
    e0, idx0 = wide.conv0_fwd(Xd, roll, w.to(DEV), b.to(DEV))
    o = wide.from_planar(e0.cpu())
    assert (o - pr).abs().max().item() <= 6e-3 * pr.abs().max().item() + 1e-6
    idx = wide.from_planar(idx0.cpu()).long()
    assert (idx[pr <= 0] == 4).all() and (idx[pr > 0] <= 3).all()
    # weight gradient from a pooled gradient + the kernel's own arg-max bytes
    de = _bf(_rand(B, C0, 32, 32, seed=5, scale=1.0 / B))
    dy = torch.zeros(B, C0, 64, 64)
    for pos in range(4):
        dy[:, :, pos // 2::2, pos % 2::2] = torch.where(idx == pos, de, torch.zeros(()))
    dw_ref = torch.nn.grad.conv2d_weight(x.double(), (C0, 3, 3, 3), dy.double(), padding=1).float()
    db_ref = dy.double().sum((0, 2, 3)).float()
    dw, db = torch.ones(C0, 3, 3, 3, device=DEV), torch.full((C0,), 2.0, device=DEV)
    wide.conv0_wgrad(Xd, torch.tensor([roll], dtype=torch.int32, device=DEV), wide.to_planar(de).to(DEV), idx0, dw, db)
    assert (dw.cpu() - 1.0 - dw_ref).abs().max().item() <= 3e-5 * dw_ref.abs().max().item() + 2e-6
    assert (db.cpu() - 2.0 - db_ref).abs().max().item() <= 3e-5 * db_ref.abs().max().item() + 2e-6


@pytest.mark.parametrize("M,N,K", [(256, 160, 1280), (37, 64, 256), (300, 1280, 160)])
def test_wide_gemm_layouts(M, N, K):
    from cgs_b200 import wide
    A, Bm, bias = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
    ref = A @ Bm.t() + bias
    tol = 2e-3 * ref.abs().max().item()                                             # TF32 operands
    out = wide.gemm(A.to(DEV), True, Bm.to(DEV), True, M, N, K, bias=bias.to(DEV))
    assert (out.cpu() - ref).abs().max().item() <= tol
    for splits in (2, 8):                                # split-K: partial tiles summed in fixed order by the last CTA of a tile
        o2 = wide.gemm(A.to(DEV), True, Bm.to(DEV), True, M, N, K, bias=bias.to(DEV), splits=splits)
        o3 = wide.gemm(A.to(DEV), True, Bm.to(DEV), True, M, N, K, bias=bias.to(DEV), splits=splits)
        assert (o2.cpu() - ref).abs().max().item() <= tol and torch.equal(o2, o3)
    out = wide.gemm(A.to(DEV), True, Bm.t().contiguous().to(DEV), False, M, N, K, bias=bias.to(DEV), relu=True)
    assert (out.cpu() - F.relu(ref)).abs().max().item() <= tol
    gate = _rand(M, N, seed=4)
    acc = torch.ones(M, N, device=DEV)
    wide.gemm(A.t().contiguous().to(DEV) if M % 4 == 0 else A.to(DEV), M % 4 != 0, Bm.t().contiguous().to(DEV), False, M, N, K, out=acc,
              gate=gate.to(DEV), accumulate=True)
    exp = 1.0 + torch.where(gate > 0, A @ Bm.t(), torch.zeros(()))
    assert (acc.cpu() - exp).abs().max().item() <= tol


def _wide_case(K, B, p, seed):
    import cgs_b200.synth as synth
    from helpers import drop_masks
    csd = synth.perturbed_state(synth.critic_shapes(K), seed, 1.5)
    X, Y, _ = synth.synthetic_frames(B, seed=seed)
    masks = drop_masks(np.random.default_rng(seed), B, K, p)
    return csd, X, Y[1, :B].astype(np.float32), masks


def _wide_oracle(csd, X, y, masks, roll, bce, q):
    from oracle import torch_ref
    sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in csd.items()}
    x = torch_ref.to_input(np.roll(X, -roll, axis=2))
    yt = torch.from_numpy(y)
    if bce:
        yt = (yt > 0.5).float()
    pred = torch_ref.critic_forward(sd, x, masks=None if masks is None else tuple(torch.from_numpy(m) for m in masks), q=q).squeeze()
    loss = F.binary_cross_entropy(pred, yt) if bce else F.mse_loss(pred, yt)
    loss.backward()
    return loss.item(), pred.detach().numpy(), {k: v.grad.numpy() for k, v in sd.items()}


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("K,B,roll,p,bce", [(5, 3, 0, 0.0, False), (5, 37, -5, 0.3, False), (2, 19, 3, 0.3, True), (5, 150, 9, 0.3, False), (4, 8, 0, 0.5, False)])
def test_wide_critic_step_vs_oracle(K, B, roll, p, bce):
    """One whole wide step (loss, predictions, all 14 parameter gradients) against (1) the oracle at the kernels' operand
    precision (bf16 operands in the 3x3 convolutions; the oracle keeps the back-propagated gradients in fp32 where the kernels
    store them as bf16 planes between layers, which is the 1-2 % that remains) and (2) the reference arithmetic."""
    from cgs_b200 import wide, ops
    from cgs_b200.nets import NewCritic
    from helpers import nhwc_masks
    from oracle import torch_ref
    csd, X, y, masks = _wide_case(K, B, p, seed=10 * K + B)
    ops.set_precision("tf32")
    try:
        c = NewCritic(chfak=K, dropout=p)
        c.load_state_dict({k: torch.from_numpy(v) for k, v in csd.items()})
        c.to(DEV).train()
        assert wide.supported(c)
        yt = torch.from_numpy(y if not bce else (y > 0.5).astype(np.float32)).to(DEV)
        loss, pred = wide.critic_train_wide(c, torch.from_numpy(X).to(DEV), yt, roll, nhwc_masks(masks, DEV), bce=bce)
        torch.cuda.synchronize()
        assert wide.status_ok()
    finally:
        ops.set_precision("fp32")
    pred = pred.cpu().numpy()
    grads = {k: v.grad.cpu().numpy() for k, v in c.named_parameters()}
    g_all = np.concatenate([grads[k].ravel() for k in grads])
    for tag, q, t_pred, t_loss, t_tot, t_one in (("operand-precision oracle", torch_ref.quant_bf16, 2e-3, 4e-3, 3e-2, 5e-2),
                                                 ("fp32 oracle", None, 1.5e-2, 3e-2, 1.5e-1, None)):
        loss_r, pred_r, grads_r = _wide_oracle(csd, X, y, masks, roll, bce, q)
        assert np.abs(pred - pred_r).max() <= t_pred, (tag, np.abs(pred - pred_r).max())
        assert abs(loss.item() - loss_r) <= t_loss * abs(loss_r) + 1e-6, (tag, loss.item(), loss_r)
        errs = {k: _rel(grads[k], grads_r[k]) for k in grads}
        tot = _rel(g_all, np.concatenate([grads_r[k].ravel() for k in grads]))
        assert tot <= t_tot, (tag, tot, errs)
        if t_one is not None:            # crit.4.bias is ONE number, a sum with cancellation: it is covered by the total
            assert max(v for k, v in errs.items() if grads[k].size > 1) <= t_one, (tag, errs)


def test_wide_critic_step_vs_reference_golden():
    """The chfak-5 critic step of the UNMODIFIED reference (tests/golden/step_c5_b2.npz) through the wide path, bf16 tolerances."""
    from cgs_b200 import wide, ops
    from cgs_b200.nets import NewCritic
    from helpers import load_golden, step_case, nhwc_masks, sample, tsd
    d = load_golden("step_c5_b2.npz")
    c = step_case(d)
    ops.set_precision("tf32")
    try:
        critic = NewCritic(chfak=c["K"], dropout=c["p"])
        critic.load_state_dict(tsd(c["csd"]))
        critic.to(DEV).train()
        X = (c["A"].permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().to(DEV)
        loss, pred = wide.critic_train_wide(critic, X, c["Y"].to(DEV), 0, nhwc_masks(c["masks"][0], DEV))
        torch.cuda.synchronize()
    finally:
        ops.set_precision("fp32")
    assert abs(loss.item() - float(d["cstep.loss"])) <= 3e-2 * abs(float(d["cstep.loss"])) + 1e-6
    assert np.abs(pred.cpu().numpy() - d["cstep.pred"]).max() <= 1.5e-2
    num = den = 0.0
    for k, v in critic.named_parameters():
        gs, ga = v.grad.double().sum().item(), float(d["cstep.gabs." + k])
        assert abs(gs - float(d["cstep.gsum." + k])) <= 5e-2 * ga + 1e-7, (k, gs, float(d["cstep.gsum." + k]), ga)
        a, b = sample(v.grad.cpu().numpy()).astype(np.float64), d["cstep.g." + k].astype(np.float64)
        num += ((a - b) ** 2).sum(); den += (b ** 2).sum()
    assert (num / den) ** 0.5 <= 0.2, (num / den) ** 0.5           # B = 2 frames: arg-max flips weigh heavily (see the oracle test for the bound at B >= 37)


def test_wide_step_is_reproducible_and_handler_takes_it():
    """Two runs give bit-identical gradients (fixed-order sums everywhere); Handler.critic_step at chfak 5 goes through the wide path
    and its loss follows the per-layer path's."""
    from cgs_b200 import wide, ops
    from cgs_b200.train_handler import Handler, parse_args
    import cgs_b200.synth as synth
    ops.set_precision("tf32")
    try:
        X, Y, _ = synth.synthetic_frames(64, seed=3)
        out = []
        for use_wide in (True, True, False):
            torch.manual_seed(5)
            H = Handler(parse_args(["--dropout", "0", "--chfak", "5"]), device=DEV)
            H.wide_critic_step = use_wide
            H.critic.to(DEV).train()
            opti = H._opt(H.critic.parameters())
            with ops.profile_calls() as prof:
                losses = [H.critic_step(torch.from_numpy(X), torch.from_numpy(Y[1, :64]).float(), opti, roll=2).item() for _ in range(3)]
            names = set(prof.summary())
            assert ("cgs_wide_conv3x3" in names) == use_wide, names
            out.append((losses, torch.cat([q.detach().reshape(-1) for q in H.critic.parameters()]).cpu()))
        assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1])
        assert np.allclose(out[0][0], out[2][0], rtol=3e-2), (out[0][0], out[2][0])
    finally:
        ops.set_precision("fp32")
