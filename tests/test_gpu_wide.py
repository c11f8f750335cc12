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
    # ReLU + pool + arg-max (+ dropout mask)
    g = torch.Generator().manual_seed(4)
    mask = ((torch.rand(B, H // 2, W // 2, Cout, generator=g) >= 0.3).float() / 0.7)
    o, idx, f32 = wide.conv3x3(xp, w.to(DEV), b.to(DEV), epi=wide.EPI_RELU_POOL, mask=mask.to(DEV), want_f32=True)
    assert wide.status_ok()
    pr, pi = F.max_pool2d(F.relu(ref), 2, return_indices=True)
    o_ref = pr * mask.permute(0, 3, 1, 2)
    o, idx = wide.from_planar(o.cpu()), wide.from_planar(idx.cpu()).long()
    assert (o - o_ref).abs().max().item() <= 6e-3 * o_ref.abs().max().item() + 1e-6
    assert (f32.cpu().permute(0, 3, 1, 2) - o_ref).abs().max().item() <= 1e-4 * o_ref.abs().max().item() + 1e-6
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
