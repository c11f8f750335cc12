"""GPU: drop-in boundary proof with the REAL reference.

The unmodified reference `main.py` (byte-identical copy in the git-ignored `oracle/_ref/`, made by `oracle/make_ref.py`;
`/root/reference` in the authoring container) is imported through `oracle/ref_shims.py`; the ONE change INTEGRATION.md §1
describes is applied — `NewCritic` / `UnetDecoder` in `main`'s namespace (what `from nets import *`, main.py:10, binds) are
the cgs_b200 classes — and then the reference's OWN `Handler.critic_pipe`, `Handler.segmentation_training` and the loop body of
`Handler.segment` run verbatim: its DataLoader-shaped batches, its `T.optim.Adam`, its `F.mse_loss` / `F.l1_loss`,
`loss.backward()`, `.item()`, `state_dict` save/load.  Results are compared with what the same loops produced on the reference's
own classes (tests/golden/loops_c1.npz).  Skipped when oracle/_ref is absent (run __graft_entry__.build() where
/root/reference exists)."""
import os

import numpy as np
import pytest
import torch

from helpers import load_golden
import cgs_b200.synth as synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("reference sources not present (oracle/_ref missing)")
    nets, main = ref_shims.load()
    import cgs_b200.nets as cn
    saved = (main.NewCritic, main.UnetDecoder)
    main.NewCritic, main.UnetDecoder = cn.NewCritic, cn.UnetDecoder          # the INTEGRATION.md §1 swap
    yield ref_shims, main
    main.NewCritic, main.UnetDecoder = saved


class _Rec:
    """Stands in for `F` inside the reference loops to record the loss terms it computes (as make_golden.py does)."""

    def __init__(self, F):
        self._F, self.log = F, {"mse_loss": [], "l1_loss": []}

    def __getattr__(self, k):
        return getattr(self._F, k)

    def mse_loss(self, a, b):
        v = self._F.mse_loss(a, b)
        self.log["mse_loss"].append(float(v.detach()))
        return v

    def l1_loss(self, a, b):
        v = self._F.l1_loss(a, b)
        self.log["l1_loss"].append(float(v.detach()))
        return v


def _handler(ref_shims, main, work, extra=()):
    import cgs_b200.nets as cn
    H = ref_shims.make_handler(["-train", "--dropout", "0", "--shift", "0", "--cepochs", "2", "--model", "g", "--saveevery", "100",
                                "--visevery", "1000000"] + list(extra), work)
    assert isinstance(H.critic, cn.NewCritic) and isinstance(H.masker, cn.UnetDecoder), "the swap did not take"
    assert H.device == "cuda"
    H.args.cload = False
    os.makedirs(os.path.join(work, H.path), exist_ok=True)       # the loops write their logs under `self.path` (main.py:94, 289)
    return H


def test_reference_critic_pipe_on_cgs_classes(ref, tmp_path):
    """Reference Handler.critic_pipe (main.py:158-236), two epochs (188 steps), exact fp32 kernels: the recorded losses follow
    the curve the reference's own classes produced, step by step, while the trajectory is still deterministic."""
    ref_shims, main = ref
    import torch.nn.functional as F
    d = load_golden("loops_c1.npz")
    N = 6000
    X, Y, I = synth.synthetic_frames(N, seed=0)
    H = _handler(ref_shims, main, str(tmp_path))
    H.critic.load_state_dict({k[len("init.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("init.c.")})
    H.X, H.Y, H.I = X, Y, I
    Xt, Yt, It = torch.from_numpy(X), torch.from_numpy(Y).t(), torch.arange(N, dtype=torch.int32)
    H.train_loader = [(Xt[i:i + 64], Yt[i:i + 64], It[i:i + 64]) for i in range(0, N, 64)]
    rec = _Rec(F)
    main.F = rec
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        H.critic_pipe(mode="train")
    finally:
        os.chdir(cwd)
        main.F = F
    ours, theirs = np.array(rec.log["mse_loss"]), d["closs"]
    assert len(ours) == 2 * 94
    assert np.abs(ours[:40] - theirs[:40]).max() <= 0.01 * theirs[:40].max()
    assert abs(np.median(ours[:94]) - np.median(theirs[:94])) <= 0.01 * np.median(theirs[:94])
    assert np.isfinite(ours).all() and ours[-20:].mean() < ours[:20].mean()
    # the reference's save/load round trip on the swapped classes (state_dict layout, main.py:136-156)
    os.chdir(str(tmp_path))
    try:
        H.save_models([H.criticname])
        sd = torch.load(H.save_paths[H.criticname], map_location="cpu")
    finally:
        os.chdir(cwd)
    assert list(sd.keys()) == [k[len("init.c."):] for k in d.files if k.startswith("init.c.")]


def test_reference_segmentation_training_and_segment_on_cgs_classes(ref, tmp_path):
    """Reference Handler.segmentation_training (main.py:314-575; -frozen, inject, L1) from the reference-trained critic, then the
    loop body of Handler.segment (main.py:1139-1164): loss curves within 1 % of the reference's own, same pos/neg split, and the
    resulting masks within 2e-2 / IoU >= 0.99 of the masks the reference's classes give with the SAME trained weights."""
    ref_shims, main = ref
    import torch.nn.functional as F
    d = load_golden("loops_c1.npz")
    N = 6000
    X, Y, I = synth.synthetic_frames(N, seed=0)
    H = _handler(ref_shims, main, str(tmp_path))
    H.critic.load_state_dict({k[len("trained.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.c.")})
    H.masker.load_state_dict({k[len("init.m."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("init.m.")})
    H.critic.to(H.device); H.masker.to(H.device)
    H.X, H.Y, H.I = X, Y, I
    H.args.frozen, H.args.live = True, False
    rec = _Rec(F)
    main.F = rec
    np.random.seed(0)
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        H.segmentation_training()
    finally:
        os.chdir(cwd)
        main.F = F
    assert len(H.Xpos) == int(d["n_pos"]) and len(H.Xneg) == int(d["n_neg"])
    mse = np.array(rec.log["mse_loss"]).reshape(-1, 2)
    l1 = np.array(rec.log["l1_loss"]) * 0.5
    assert len(l1) == len(d["seg_l1"])
    sm = lambda v: np.convolve(v, np.ones(30) / 30, mode="valid")
    rel = np.abs(sm(l1) - sm(d["seg_l1"])) / sm(d["seg_l1"])
    assert rel.max() < 0.01, rel.max()
    ours, theirs = sm(mse.sum(1)), sm(d["seg_replace"] + d["seg_inject"])
    assert np.abs(ours - theirs).max() <= 0.01 * theirs.max() + 1e-6
    # ---- the loop body of Handler.segment (main.py:1134-1151, 1164), verbatim, on the classes as the reference left them
    H.masker.load_state_dict({k[len("trained.m."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.m.")})
    H.critic.eval(); H.masker.eval()
    Xs = X[:32] / 255.0
    batch = torch.from_numpy(Xs).permute(0, 3, 1, 2).float().to(H.device)
    pred, embeds = H.critic(batch, collect=True)
    mask = H.masker(batch, embeds).detach().cpu().numpy()
    hard = mask >= 0.1
    ref_hard = np.unpackbits(d["proc_hard"])[:hard.size].reshape(hard.shape).astype(bool)
    assert np.abs(mask - d["proc_mask"]).max() <= 2e-2
    assert (hard | ref_hard).sum() == 0 or (hard & ref_hard).sum() / (hard | ref_hard).sum() >= 0.99
    assert np.abs(pred.detach().cpu().numpy() - d["proc_pred"]).max() <= 5e-3
