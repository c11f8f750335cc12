"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The reference classes (nets.NewCritic,
nets.UnetDecoder, main.Handler) are imported via oracle/ref_shims.py and executed
by the installed torch on CPU in fp32; nothing here is product code.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
import cgs_b200.synth as synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
nets, main = ref_shims.load()
torch.set_num_threads(8)


def np_sd(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def load_np(module, sd):
    module.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})


def drop_masks(rng, B, c, p):
    """Shared dropout masks (multiplicative, 0 or 1/(1-p)), logical NCHW."""
    if p <= 0:
        return None
    mk = lambda *s: ((rng.random(s) >= p).astype(np.float32) / np.float32(1 - p))
    return (mk(B, 8 * c, 8, 8), mk(B, 16 * c, 4, 4), mk(B, 32 * c))


class MaskedCritic:
    """Runs the reference NewCritic in train mode with injected dropout masks."""

    def __init__(self, critic):
        self.c = critic

    def __call__(self, x, masks, collect=False):
        c = self.c
        if masks is None:
            c.eval()
            return c(x, collect=collect)
        c.train()
        f9, f13, c3 = c.features[9], c.features[13], c.crit[3]
        m = [torch.from_numpy(a) for a in masks]
        f9.forward = lambda t: t * m[0]
        f13.forward = lambda t: t * m[1]
        c3.forward = lambda t: t * m[2]
        try:
            return c(x, collect=collect)
        finally:
            del f9.forward, f13.forward, c3.forward


def sample(a, n=257):
    a = np.asarray(a).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(n, a.size)).astype(np.int64)
    return a[idx]


def gen_init_kat():
    """§8c KAT: seed-0 default init of the reference modules."""
    for K in (1, 5):
        torch.manual_seed(0)
        c = nets.NewCritic(bottleneck=32, chfak=K, dropout=0.3).eval()
        m = nets.UnetDecoder(bottleneck=32, chfak=K).eval()
        x = torch.rand(4, 3, 64, 64)
        p, e = c(x, collect=True)
        z = m(x, e)
        out = dict(pred=p.detach().numpy(), mask_min=z.min().item(), mask_max=z.max().item(),
                   mask_mean=z.double().mean().item(), ge_half=int((z >= 0.5).sum()),
                   embed_sums=np.array([t.double().sum().item() for t in e]),
                   z_row=z[0, 0, 0, :8].detach().numpy(), x=sample(x.numpy()))
        if K == 1:
            out.update({"c." + k: v for k, v in np_sd(c).items()})
            out.update({"m." + k: v for k, v in np_sd(m).items()})
            out["mask"] = z.detach().numpy()
            out["x_full"] = x.numpy()
        np.savez_compressed(f"{OUT}/kat_init_c{K}.npz", **out)
        print("kat", K, p.flatten().tolist(), out["ge_half"])


def gen_step(K, B, p, seed, scale, full):
    """Forward / backward goldens on perturbed weights: critic step, Hourglass step
    (live+inject+L1+L2 and frozen variants), saliency input-grad, inference."""
    rng = np.random.default_rng(1000 + seed)
    csd = synth.perturbed_state(synth.critic_shapes(K), seed, scale)
    msd = synth.perturbed_state(synth.masker_shapes(K), seed + 1, scale)
    critic = nets.NewCritic(bottleneck=32, chfak=K, dropout=p)
    masker = nets.UnetDecoder(bottleneck=32, chfak=K)
    load_np(critic, csd)
    load_np(masker, msd)
    X, Y, _ = synth.synthetic_frames(2 * B, seed=seed)
    A = torch.from_numpy(X[:B]).permute(0, 3, 1, 2).float() / 255.0
    Bf = torch.from_numpy(X[B:]).permute(0, 3, 1, 2).float() / 255.0
    Yt = torch.from_numpy(Y[1, :B]).float()
    masks = [drop_masks(rng, B, K, p) for _ in range(4)]
    mc = MaskedCritic(critic)
    out = dict(K=K, B=B, p=p, seed=seed, scale=scale)
    keep = (lambda a: np.asarray(a)) if full else sample

    # ---- eval-mode inference (Handler.segment body, main.py:1139-1151,1164)
    critic.eval(); masker.eval()
    pred, embeds = critic(A, collect=True)
    mask = masker(A, embeds)
    out["inf.pred"] = pred.detach().numpy()
    out["inf.mask"] = keep(mask.detach().numpy())
    out["inf.mask_sum"] = mask.double().sum().item()
    out["inf.hard_count"] = np.array([int((mask >= t).sum()) for t in (0.1, 0.3, 0.5, 0.7)])
    for i, e in enumerate(embeds):
        out[f"inf.e{i}"] = keep(e.detach().numpy())
        out[f"inf.e{i}_sum"] = e.double().sum().item()

    # ---- saliency (main.py:1137-1148): d mean(pred) / d batch
    a = A.clone().requires_grad_(True)
    critic(a).mean().backward()
    out["sal.grad_abs_sum"] = keep(a.grad.abs().sum(1).numpy())
    out["sal.total"] = a.grad.double().abs().sum().item()

    # ---- critic_pipe step (main.py:191-198), train mode with shared masks
    critic.zero_grad()
    pr = mc(A, masks[0]).squeeze()
    loss = F.mse_loss(pr, Yt)
    loss.backward()
    out["cstep.loss"] = loss.item()
    out["cstep.pred"] = pr.detach().numpy()
    for k, v in critic.named_parameters():
        out["cstep.g." + k] = keep(v.grad.numpy())
        out["cstep.gsum." + k] = v.grad.double().sum().item()
        out["cstep.gabs." + k] = v.grad.double().abs().sum().item()
    critic.zero_grad()
    pr = mc(A, masks[0]).squeeze()
    bl = F.binary_cross_entropy(pr, (Yt > 0.5).float())      # --threshrew variant, main.py:192-193
    bl.backward()
    out["cstep.bce"] = bl.item()
    out["cstep.bce.gsum"] = np.array([v.grad.double().sum().item() for v in critic.parameters()])
    out["cstep.bce.gabs"] = np.array([v.grad.double().abs().sum().item() for v in critic.parameters()])

    # ---- segmentation_training step (main.py:364-462)
    for tag, live, inject, L1, L2, static in (("hg_full", True, True, 0.5, 0.25, True),
                                              ("hg_frozen", False, True, 0.5, 0.0, True),
                                              ("hg_noinj", True, False, 0.0, 0.5, False)):
        critic.zero_grad(); masker.zero_grad()
        masker.train()
        pred, embeds = mc(A, masks[0], collect=True)
        negpred = mc(Bf, masks[1])
        pred = pred.squeeze(); negpred = negpred.squeeze().detach()
        loss = 0
        if live:
            cl = F.mse_loss(pred, Yt); loss = loss + 5 * cl; out[f"{tag}.critic"] = cl.item()
        Z = masker(A, embeds)
        replaced = A * (1 - Z) + Z * Bf
        rl = F.mse_loss(mc(replaced, masks[2]).squeeze(), negpred.detach()); loss = loss + rl
        out[f"{tag}.replace"] = rl.item()
        if inject:
            injected = Bf * (1 - Z) + Z * A
            il = F.mse_loss(mc(injected, masks[3]).squeeze(), pred.detach()); loss = loss + il
            out[f"{tag}.inject"] = il.item()
        vf = 1 if static else 1 - pred.detach().view(-1, 1, 1, 1)
        if L1:
            n1 = L1 * F.l1_loss(vf * Z, torch.zeros_like(Z)); loss = loss + n1; out[f"{tag}.L1"] = n1.item()
        if L2:
            n2 = L2 * F.mse_loss(vf * Z, torch.zeros_like(Z)); loss = loss + n2; out[f"{tag}.L2"] = n2.item()
        loss.backward()
        out[f"{tag}.loss"] = loss.item()
        out[f"{tag}.Z"] = keep(Z.detach().numpy())
        for pre, mod in (("c", critic), ("m", masker)):
            for k, v in mod.named_parameters():
                g = v.grad if v.grad is not None else torch.zeros_like(v)
                out[f"{tag}.g.{pre}.{k}"] = keep(g.numpy())
                out[f"{tag}.gsum.{pre}.{k}"] = g.double().sum().item()
                out[f"{tag}.gabs.{pre}.{k}"] = g.double().abs().sum().item()
    name = f"{OUT}/step_c{K}_b{B}.npz"
    np.savez_compressed(name, **out)
    print("step", K, B, {k: out[k] for k in out if k.endswith(".loss")}, os.path.getsize(name))


class _Rec:
    """Proxy for torch.nn.functional inside main.py that records loss values."""

    def __init__(self):
        self.log = {"mse_loss": [], "l1_loss": []}

    def __getattr__(self, name):
        fn = getattr(F, name)
        if name in self.log:
            def wrapped(*a, **k):
                r = fn(*a, **k)
                self.log[name].append(r.item())
                return r
            return wrapped
        return fn


def gen_loops():
    """Loop-level goldens: the reference Handler loops run verbatim (dropout 0, shift 0,
    fixed batch order) on the synthetic set; loss curves + trained weights + -process masks."""
    work = "/tmp/cgs_golden_work"
    N = 6000
    X, Y, I = synth.synthetic_frames(N, seed=0)
    H = ref_shims.make_handler(["-train", "--dropout", "0", "--shift", "0", "--cepochs", "11",
                                "--model", "g", "--saveevery", "100", "--visevery", "1000000"], work)
    torch.manual_seed(0)
    np.random.seed(0)
    H.reset_models()
    H.models[H.criticname] = H.critic
    H.models[H.maskername] = H.masker
    H.args.cload = False
    init_c, init_m = np_sd(H.critic), np_sd(H.masker)
    H.X, H.Y, H.I = X, Y, I
    bs = 64
    Xt, Yt, It = torch.from_numpy(X), torch.from_numpy(Y).t(), torch.arange(N, dtype=torch.int32)
    H.train_loader = [(Xt[i:i + bs], Yt[i:i + bs], It[i:i + bs]) for i in range(0, N, bs)]
    rec = _Rec()
    main.F = rec
    cwd = os.getcwd()
    os.chdir(work)
    try:
        H.critic_pipe(mode="train")
        closs = np.array(rec.log["mse_loss"], dtype=np.float64)
        trained_c = np_sd(H.critic)
        rec.log = {"mse_loss": [], "l1_loss": []}
        H.args.frozen, H.args.live = True, False
        H.segmentation_training()
        mse = np.array(rec.log["mse_loss"], dtype=np.float64).reshape(-1, 2)   # replace, inject per step
        l1 = np.array(rec.log["l1_loss"], dtype=np.float64)
        trained_m = np_sd(H.masker)
    finally:
        os.chdir(cwd)
        main.F = F
    print("critic steps", len(closs), closs[:3], closs[-3:], "seg steps", len(l1), mse[-1], l1[-1])
    # -process on a slice (main.py:1130-1164), eval mode, float64 /255 then .float()
    H.critic.eval(); H.masker.eval()
    Xs = X[:32] / 255.0
    batch = torch.from_numpy(Xs).permute(0, 3, 1, 2).float()
    pred, embeds = H.critic(batch, collect=True)
    mask = H.masker(batch, embeds).detach().numpy()
    out = dict(closs=closs, seg_replace=mse[:, 0], seg_inject=mse[:, 1], seg_l1=l1 * 0.5,
               n_pos=len(H.Xpos), n_neg=len(H.Xneg), proc_pred=pred.detach().numpy(),
               proc_mask=mask.astype(np.float32), proc_hard=np.packbits(mask >= 0.1))
    for pre, sd in (("init.c.", init_c), ("init.m.", init_m), ("trained.c.", trained_c), ("trained.m.", trained_m)):
        out.update({pre + k: v for k, v in sd.items()})
    np.savez_compressed(f"{OUT}/loops_c1.npz", **out)
    print("coverage@0.1", (mask >= 0.1).mean(), "size", os.path.getsize(f"{OUT}/loops_c1.npz"))


if __name__ == "__main__":
    which = sys.argv[1:] or ["kat", "step", "loops"]
    if "kat" in which:
        gen_init_kat()
    if "step" in which:
        gen_step(K=1, B=6, p=0.3, seed=11, scale=1.6, full=True)
        gen_step(K=2, B=3, p=0.5, seed=12, scale=1.6, full=False)
        gen_step(K=5, B=2, p=0.5, seed=13, scale=1.5, full=False)
    if "loops" in which:
        gen_loops()


def gen_envelope():
    """Sensitivity envelope of the critic loss curve: the loop of gen_loops() re-run by the oracle
    (bit-identical to the reference on this curve, see tests/test_oracle.py) from initial weights
    perturbed by 1e-6 relative (fp32-rounding scale).  Training is chaotic in the fast-learning
    phase, so an implementation can only be asked to stay inside this envelope."""
    from oracle import torch_ref
    d = np.load(f"{OUT}/loops_c1.npz")
    X, Y, _ = synth.synthetic_frames(6000, seed=0)
    Xt, Yt = torch.from_numpy(X), torch.from_numpy(Y).t()

    def run(eps, seed):
        g = torch.Generator().manual_seed(seed)
        sd = {k[len("init.c."):]: torch.from_numpy(d[k]).clone() for k in d.files if k.startswith("init.c.")}
        for v in sd.values():
            v.mul_(1 + eps * (torch.rand(v.shape, generator=g) - 0.5))
            v.requires_grad_(True)
        opt = torch.optim.Adam(sd.values())
        out = []
        for ep in range(11):
            for i in range(0, 6000, 64):
                loss, _ = torch_ref.critic_loss(sd, Xt[i:i + 64].permute(0, 3, 1, 2).float() / 255.0, Yt[i:i + 64, 1].float())
                opt.zero_grad(); loss.backward(); opt.step()
                out.append(loss.item())
        return np.array(out)
    sm = lambda v: np.convolve(v, np.ones(30) / 30, mode="valid")
    ref = d["closs"]
    assert np.array_equal(run(0.0, 0), ref), "oracle loop no longer reproduces the reference curve"
    devs = [np.abs(sm(run(1e-6, s)) - sm(ref)) / sm(ref) for s in (1, 2, 3, 4)]
    env = np.max(devs, axis=0)
    np.savez_compressed(f"{OUT}/loops_envelope_c1.npz", env=env.astype(np.float32), eps=1e-6, window=30)
    print("envelope max", env.max(), "median", np.median(env), "last", env[-1])


def gen_envelope_epochs(seeds=(1, 2, 3, 4, 5, 6, 7, 8)):
    """Per-epoch band of the critic loss curve: the reference loop (oracle, bit-identical to the reference on this curve) re-run
    from initial weights perturbed by 1e-6 relative, 8 seeds.  Stored: the per-epoch median loss of every run (94 steps per epoch,
    11 epochs) -> tests require an implementation's per-epoch medians to sit inside [min, max] over the runs (widened as the test
    states).  VERDICT round 1, next #8."""
    from oracle import torch_ref
    d = np.load(f"{OUT}/loops_c1.npz")
    X, Y, _ = synth.synthetic_frames(6000, seed=0)
    Xt, Yt = torch.from_numpy(X), torch.from_numpy(Y).t()

    def run(eps, seed):
        g = torch.Generator().manual_seed(seed)
        sd = {k[len("init.c."):]: torch.from_numpy(d[k]).clone() for k in d.files if k.startswith("init.c.")}
        for v in sd.values():
            v.mul_(1 + eps * (torch.rand(v.shape, generator=g) - 0.5))
            v.requires_grad_(True)
        opt = torch.optim.Adam(sd.values())
        out = []
        for ep in range(11):
            for i in range(0, 6000, 64):
                loss, _ = torch_ref.critic_loss(sd, Xt[i:i + 64].permute(0, 3, 1, 2).float() / 255.0, Yt[i:i + 64, 1].float())
                opt.zero_grad(); loss.backward(); opt.step()
                out.append(loss.item())
        return np.array(out)
    ep = lambda v: np.array([np.median(v[i:i + 94]) for i in range(0, 1034, 94)])
    path = f"{OUT}/loops_envelope_epochs_c1.npz"
    if os.path.exists(path) and "epoch_medians" in np.load(path).files:
        med = np.load(path)["epoch_medians"]                      # the 1e-6 band is kept; only the operand-precision band is added
    else:
        med = np.stack([ep(d["closs"])] + [ep(run(1e-6, s)) for s in seeds])
    # the same loop perturbed at bf16 operand precision (2^-9 relative): the yardstick for an implementation whose convolution
    # operands are rounded to bf16 at every step (the onset of the fast-learning phase moves by more than under 1e-6)
    med_bf = np.stack([ep(run(2.0 ** -9, 100 + s)) for s in seeds])
    np.savez_compressed(path, epoch_medians=med.astype(np.float64), eps=1e-6, seeds=np.array((0,) + tuple(seeds)),
                        epoch_medians_bf16=med_bf.astype(np.float64), eps_bf16=2.0 ** -9)
    print("per-epoch min", med.min(0)); print("per-epoch max", med.max(0)); print("max/min", med.max(0) / med.min(0))
    print("bf16-precision band: min", med_bf.min(0)); print("max", med_bf.max(0)); print("max/min", med_bf.max(0) / med_bf.min(0))


if __name__ == "__main__" and "envelope_epochs" in sys.argv[1:]:
    gen_envelope_epochs()
    sys.exit(0)

if __name__ == "__main__" and "envelope" in sys.argv[1:]:
    gen_envelope()
