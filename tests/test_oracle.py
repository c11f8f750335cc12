"""CPU: the oracle restatement (oracle/torch_ref.py) against the golden vectors produced by the
unmodified reference classes (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import torch_ref
from helpers import assert_close, load_golden, sample, step_case, tmasks, tsd

STEP_FILES = ["step_c1_b6.npz", "step_c2_b3.npz", "step_c5_b2.npz"]


def test_init_kat_c1():
    d = load_golden("kat_init_c1.npz")
    csd = {k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("c.")}
    msd = {k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("m.")}
    x = torch.from_numpy(d["x_full"])
    pred, embeds = torch_ref.critic_forward(csd, x, collect=True)
    z = torch_ref.decoder_forward(msd, x, embeds)
    # SURVEY.md §8c KAT numbers (seed-0 reference init)
    np.testing.assert_allclose(pred.flatten().numpy(), [0.46741211, 0.46733087, 0.46730250, 0.46742359], rtol=0, atol=1e-7)
    assert int((z >= 0.5).sum()) == 16263 == int(d["ge_half"])
    np.testing.assert_array_equal(z.numpy(), d["mask"])
    np.testing.assert_array_equal(pred.numpy(), d["pred"])
    np.testing.assert_allclose([e.double().sum().item() for e in embeds], d["embed_sums"], rtol=1e-12)


@pytest.mark.parametrize("fname", STEP_FILES)
def test_step_goldens(fname):
    d = load_golden(fname)
    c = step_case(d)
    full = c["K"] == 1
    keep = (lambda a: np.asarray(a)) if full else sample
    csd, msd = tsd(c["csd"]), tsd(c["msd"])
    for t in list(csd.values()) + list(msd.values()):
        t.requires_grad_(True)
    A, Bf, Y = c["A"], c["Bf"], c["Y"]
    # inference
    pred, embeds = torch_ref.critic_forward(csd, A, collect=True)
    mask = torch_ref.decoder_forward(msd, A, embeds)
    np.testing.assert_array_equal(pred.detach().numpy(), d["inf.pred"])
    np.testing.assert_array_equal(keep(mask.detach().numpy()), d["inf.mask"])
    for i, e in enumerate(embeds):
        np.testing.assert_array_equal(keep(e.detach().numpy()), d[f"inf.e{i}"])
    assert [int((mask >= t).sum()) for t in (0.1, 0.3, 0.5, 0.7)] == d["inf.hard_count"].tolist()
    # saliency
    a = A.clone().requires_grad_(True)
    torch_ref.critic_forward(csd, a).mean().backward()
    assert_close(keep(a.grad.abs().sum(1).numpy()), d["sal.grad_abs_sum"], rtol=1e-6, atol=1e-9, what="saliency")
    for t in csd.values():
        t.grad = None
    # critic step
    masks = [tmasks(m) for m in c["masks"]]
    loss, pr = torch_ref.critic_loss(csd, A, Y, masks=masks[0])
    loss.backward()
    assert abs(loss.item() - float(d["cstep.loss"])) <= 1e-7
    for k, t in csd.items():
        assert_close(keep(t.grad.numpy()), d["cstep.g." + k], rtol=1e-5, atol=1e-8, what="cstep.g." + k)
        t.grad = None
    bl, _ = torch_ref.critic_loss(csd, A, (Y > 0.5).float(), masks=masks[0], threshrew=True)
    assert abs(bl.item() - float(d["cstep.bce"])) <= 1e-6
    # hourglass variants
    for tag, live, inject, L1, L2, static in (("hg_full", True, True, 0.5, 0.25, True),
                                              ("hg_frozen", False, True, 0.5, 0.0, True),
                                              ("hg_noinj", True, False, 0.0, 0.5, False)):
        for t in list(csd.values()) + list(msd.values()):
            t.grad = None
        loss, terms, Z = torch_ref.hourglass_losses(csd, msd, A, Bf, Y, live=live, inject=inject, L1=L1, L2=L2,
                                                    staticnorm=static, masks=masks)
        loss.backward()
        assert abs(loss.item() - float(d[f"{tag}.loss"])) <= 2e-6 * max(1.0, abs(loss.item()))
        for name, val in terms.items():
            assert abs(val.item() - float(d[f"{tag}.{name}"])) <= 1e-6, (tag, name)
        np.testing.assert_allclose(keep(Z.detach().numpy()), d[f"{tag}.Z"], rtol=0, atol=1e-7)
        for pre, sd in (("c", csd), ("m", msd)):
            for k, t in sd.items():
                g = t.grad if t.grad is not None else torch.zeros_like(t)
                assert_close(keep(g.numpy()), d[f"{tag}.g.{pre}.{k}"], rtol=1e-4, atol=1e-7, what=f"{tag}.g.{pre}.{k}")


def test_loops_fixture_process_masks():
    d = load_golden("loops_c1.npz")
    csd = {k[len("trained.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.c.")}
    msd = {k[len("trained.m."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.m.")}
    import cgs_b200.synth as synth
    X, _, _ = synth.synthetic_frames(6000, seed=0)
    batch = torch.from_numpy(X[:32] / 255.0).permute(0, 3, 1, 2).float()
    pred, mask, hard = torch_ref.segment_batch(csd, msd, batch, 0.1)
    np.testing.assert_allclose(mask.numpy(), d["proc_mask"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(np.packbits(hard.numpy()), d["proc_hard"])
    assert len(d["closs"]) == 1034 and d["closs"][-1] < 0.01 < d["closs"][0]


def test_adam_restatement_matches_torch():
    torch.manual_seed(1)
    p = [torch.randn(7, 5), torch.randn(3)]
    q = [t.clone().requires_grad_(True) for t in p]
    opt = torch.optim.Adam(q)
    state = {}
    for it in range(5):
        g = [torch.randn_like(t) for t in p]
        for t, gg in zip(q, g):
            t.grad = gg.clone()
        opt.step()
        torch_ref.adam_step(p, g, state)
    for a, b in zip(p, q):
        np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=1e-6, atol=1e-7)


def test_shift_batch_matches_roll():
    X = torch.arange(2 * 4 * 8 * 3, dtype=torch.uint8).reshape(2, 4, 8, 3)
    assert torch.equal(torch_ref.shift_batch(X, 3, True), torch.roll(X, -3, dims=2))
    assert torch.equal(torch_ref.shift_batch(X, 3, False), torch.roll(X, 3, dims=2))
