"""GPU: the byte kernels either side of the hot path (csrc/edges.cu) against numpy restatements of the
reference expressions; integer / byte outputs bit-exact."""
import numpy as np
import pytest
import torch

import cgs_b200.synth as synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_gather_frames_matches_numpy_indexing():
    """`self.Xpos[Hidx]`, `self.Xneg[Lidx]`, `self.Xneg[Cidx]` (reference main.py:345-353), incl. repeated indices."""
    import cgs_b200.ops as ops
    X, _, _ = synth.synthetic_frames(700, seed=4)
    rng = np.random.default_rng(0)
    idx = rng.choice(700, 2048).astype(np.int32)
    out = ops.gather_frames(torch.from_numpy(X).to(DEV), torch.from_numpy(idx).to(DEV))
    assert np.array_equal(out.cpu().numpy(), X[idx])
    one = ops.gather_frames(torch.from_numpy(X).to(DEV), torch.tensor([699], dtype=torch.int32, device=DEV))
    assert np.array_equal(one.cpu().numpy()[0], X[699])


@pytest.mark.parametrize("B", [1, 37])
def test_mask_images_match_reference_expressions(B):
    """np.stack([X] + [np.concatenate((m, m, m), axis=1).transpose(0, 2, 3, 1) for m in (M, hardM)], axis=1), then
    (masks * 255).astype(np.uint8), per frame and as the -concatenated strip (reference main.py:1127, 1164, 1212-1223)."""
    import cgs_b200.ops as ops
    rng = np.random.default_rng(B)
    Xu8, _, _ = synth.synthetic_frames(B, seed=B)
    M = rng.random((B, 1, 64, 64)).astype(np.float32)
    M.reshape(-1)[:64] = np.linspace(0, 1, 64, dtype=np.float32)               # exact 0, 1 and k/255-ish values
    hardM = M >= 0.1
    X = Xu8 / 255.0                                                             # main.py:1127 (float64)
    masks = np.stack([X] + [np.concatenate((m, m, m), axis=1).transpose(0, 2, 3, 1) for m in (M, hardM)], axis=1)
    ref_raw = (masks[:, 1] * 255).astype(np.uint8)
    ref_thr = (masks[:, 2] * 255).astype(np.uint8)
    ref_strip = np.stack([np.concatenate((masks[f] * 255).astype(np.uint8), axis=-2) for f in range(B)])
    Md, Hd = torch.from_numpy(M).to(DEV), torch.from_numpy(hardM.astype(np.uint8)).to(DEV)
    raw, thr = ops.mask_images(Md, Hd)
    assert np.array_equal(raw.cpu().numpy(), ref_raw) and np.array_equal(thr.cpu().numpy(), ref_thr)
    strip = ops.mask_images(Md, Hd, torch.from_numpy(Xu8).to(DEV), concatenated=True)
    assert np.array_equal(strip.cpu().numpy(), ref_strip)


@pytest.mark.parametrize("thresh,glob", [(0.3, False), (0.9, False), (0.05, False), (0.3, True)])
def test_saliency_normalize_matches_reference_expressions(thresh, glob):
    """reference main.py:974-993: k-th order statistic per frame (np.sort(...)[k]) or the global mean norm, scale by pred, clip,
    threshold."""
    import sys
    import cgs_b200.ops as ops
    B = 33
    rng = np.random.default_rng(1)
    salM = np.abs(rng.standard_normal((B, 1, 64, 64))).astype(np.float32) * rng.random((B, 1, 1, 1)).astype(np.float32)
    salM[3] = 0.0                                                               # a frame without any gradient
    salM[4, 0, :32] = salM[4, 0, 0, 0]                                          # heavy ties
    preds = rng.random(B).astype(np.float32)
    s = salM.copy()
    if glob:
        norm = (s * (s >= 0)).mean() * thresh
    else:
        k = int(s.shape[-1] * s.shape[-2] * thresh)
        norm = np.sort(s.reshape(B, 1, -1), axis=-1)[:, :, k, None, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        s = s / (norm + sys.float_info.min)
        s = s * preds[:, None, None, None]
    s = s.astype(np.float32)
    s[(s >= 1)] = 1
    hard = (s > thresh).astype(np.uint8)
    out, h = ops.saliency_normalize(torch.from_numpy(salM).to(DEV), torch.from_numpy(preds).to(DEV), thresh, global_norm=glob)
    o = out.cpu().numpy()
    if glob:
        ok = np.isfinite(s)
        assert np.allclose(o[ok], s[ok], rtol=1e-5, atol=1e-7)
        assert (h.cpu().numpy() != hard).mean() < 1e-4
    else:
        assert np.array_equal(np.isnan(o), np.isnan(s))
        ok = ~np.isnan(s)
        assert np.array_equal(o[ok], s[ok])
        assert np.array_equal(h.cpu().numpy(), hard)


def test_segmentation_training_device_gather_equals_host_gather():
    """segmentation_training with the device-resident dataset + cgs_gather_frames gives the very same loss sequence as with
    the host-side numpy gather of the reference (main.py:345-353): same indices, same frames, same kernels."""
    import cgs_b200.ops as ops
    from helpers import load_golden
    from cgs_b200.train_handler import Handler, parse_args
    d = load_golden("loops_c1.npz")
    X, Y, I = synth.synthetic_frames(3000, seed=0)
    logs = []
    ops.set_precision("tf32")
    try:
        for dev_data in (True, False):
            a = parse_args(["-frozen", "--dropout", "0", "--shift", "7", "--saveevery", "100", "--model", "/tmp/cgs_gather_test"])
            a.cload = False
            torch.manual_seed(3)
            np.random.seed(3)
            H = Handler(a, device=DEV)
            H.device_dataset = dev_data
            H.critic.load_state_dict({k[len("trained.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.c.")})
            H.masker.load_state_dict({k[len("init.m."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("init.m.")})
            H.critic.to(DEV); H.masker.to(DEV)
            H.X, H.Y, H.I = X, Y, I
            H.segmentation_training()
            logs.append(np.array([[t[k] for k in sorted(t)] for t in H.seg_log]))
    finally:
        ops.set_precision("fp32")
    assert logs[0].shape == logs[1].shape and logs[0].shape[0] >= 10
    # identical frames reach identical kernels; the loss scalars are summed with float atomics (order varies run to run)
    assert np.allclose(logs[0], logs[1], rtol=2e-3, atol=1e-7), np.abs(logs[0] - logs[1]).max()
