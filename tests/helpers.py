"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

import cgs_b200.synth as synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def sample(a, n=257):
    """Must match tests/golden/make_golden.py::sample."""
    a = np.asarray(a).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(n, a.size)).astype(np.int64)
    return a[idx]


def drop_masks(rng, B, c, p):
    """Must match tests/golden/make_golden.py::drop_masks (logical NCHW)."""
    if p <= 0:
        return None
    mk = lambda *s: ((rng.random(s) >= p).astype(np.float32) / np.float32(1 - p))
    return (mk(B, 8 * c, 8, 8), mk(B, 16 * c, 4, 4), mk(B, 32 * c))


def step_case(d):
    """Rebuild the inputs of a step_c*_b*.npz fixture from its recipe."""
    K, B, p, seed, scale = int(d["K"]), int(d["B"]), float(d["p"]), int(d["seed"]), float(d["scale"])
    rng = np.random.default_rng(1000 + seed)
    csd = synth.perturbed_state(synth.critic_shapes(K), seed, scale)
    msd = synth.perturbed_state(synth.masker_shapes(K), seed + 1, scale)
    X, Y, _ = synth.synthetic_frames(2 * B, seed=seed)
    A = torch.from_numpy(X[:B]).permute(0, 3, 1, 2).float() / 255.0
    Bf = torch.from_numpy(X[B:]).permute(0, 3, 1, 2).float() / 255.0
    Yt = torch.from_numpy(Y[1, :B]).float()
    masks = [drop_masks(rng, B, K, p) for _ in range(4)]
    return dict(K=K, B=B, p=p, csd=csd, msd=msd, A=A, Bf=Bf, Y=Yt, masks=masks)


def tsd(sd, device="cpu", dtype=torch.float32):
    return {k: torch.from_numpy(np.asarray(v)).to(device=device, dtype=dtype) for k, v in sd.items()}


def tmasks(masks, device="cpu"):
    return None if masks is None else tuple(torch.from_numpy(m).to(device) for m in masks)


def nhwc_masks(masks, device):
    """Logical-NCHW numpy masks -> the NHWC device tensors NewCritic._forced_masks expects."""
    if masks is None:
        return (None, None, None)
    m0, m1, m2 = (torch.from_numpy(m).to(device) for m in masks)
    return (m0.permute(0, 2, 3, 1).contiguous(), m1.permute(0, 2, 3, 1).contiguous(), m2.contiguous())


def assert_close(a, b, rtol=1e-4, atol=1e-6, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()}/{bad.size} elements out of tolerance; "
                           f"max abs err {err.max():.3e}, max |ref| {np.abs(b).max():.3e}")
