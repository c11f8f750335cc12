"""Generate tests/golden/loops_q_c1.npz: the 94-step segmentation_training phase of tests/golden/loops_c1.npz re-run by the
ORACLE (oracle/torch_ref.py; bit-identical to the unmodified reference on this curve, asserted below)
  * at the operand precision of the whole-frame kernels: all bf16 (`q_bf16`, the default fused step), bf16 Hourglass + TF32
    scoring passes (`q_bf16_tf32`, Handler.hg_score_bf16 = False) and all TF32 (`q_tf32`), and
  * in reference arithmetic from masker weights perturbed by 1e-6 relative (4 seeds): the reference's own sensitivity
    envelope for this phase.
Authoring container only (CPU, ~1 minute):  python tests/golden/make_golden_q.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import torch_ref  # noqa: E402
import cgs_b200.synth as synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def run(d, X, Y, q_embed=None, q_score=None, q_mask=None, eps=0.0, seed=0):
    csd = {k[len("trained.c."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("trained.c.")}
    msd = {k[len("init.m."):]: torch.from_numpy(d[k]).clone() for k in d.files if k.startswith("init.m.")}
    g = torch.Generator().manual_seed(seed)
    for v in msd.values():
        if eps:
            v.mul_(1 + eps * (torch.rand(v.shape, generator=g) - 0.5))
        v.requires_grad_(True)
    with torch.no_grad():                                            # extract_contrastive_data, main.py:238-312
        preds = torch.cat([torch_ref.critic_forward(csd, torch_ref.to_input(X[i:i + 128])).squeeze(1) for i in range(0, len(X), 128)])
    pos, neg = (preds > 0.7).numpy(), (preds < 0.3).numpy()
    assert pos.sum() == int(d["n_pos"]) and neg.sum() == int(d["n_neg"])
    Xpos, Xneg = X[pos], X[neg]
    opt = torch.optim.Adam(msd.values())                             # main.py:334 (-frozen)
    np.random.seed(0)
    out = []
    for _ in range(int(np.ceil(len(Xpos) / 32))):                    # main.py:340-463, one epoch
        H, L, C = np.random.choice(len(Xpos), 32), np.random.choice(len(Xneg), 32), np.random.choice(len(Xneg), 64)
        A = torch_ref.to_input(np.concatenate((Xpos[H], Xneg[L])))
        B = torch_ref.to_input(Xneg[C])
        loss, terms, _ = torch_ref.hourglass_losses(csd, msd, A, B, None, live=False, inject=True, L1=0.5, q_embed=q_embed,
                                                    q_score=q_score, q_mask=q_mask)
        opt.zero_grad(); loss.backward(); opt.step()
        out.append([terms["replace"].item(), terms["inject"].item(), terms["L1"].item()])
    return np.array(out)


if __name__ == "__main__":
    d = np.load(f"{OUT}/loops_c1.npz")
    X, Y, _ = synth.synthetic_frames(6000, seed=0)
    ref = np.stack([d["seg_replace"], d["seg_inject"], d["seg_l1"]], axis=1)
    base = run(d, X, Y)
    print("oracle vs reference curve: max |diff|", np.abs(base - ref).max())
    assert np.allclose(base, ref, rtol=1e-5, atol=1e-8), "oracle loop no longer reproduces the reference curve"
    qrun = run(d, X, Y, q_embed=torch_ref.quant_bf16, q_score=torch_ref.quant_bf16, q_mask=torch_ref.quant_bf16)
    mix = run(d, X, Y, q_embed=torch_ref.quant_bf16, q_score=torch_ref.quant_tf32, q_mask=torch_ref.quant_bf16)
    tf = run(d, X, Y, q_embed=torch_ref.quant_tf32, q_score=torch_ref.quant_tf32, q_mask=torch_ref.quant_tf32)
    pert = np.stack([run(d, X, Y, eps=1e-6, seed=s) for s in (1, 2, 3, 4)])
    sm = lambda v: np.convolve(v, np.ones(30) / 30, mode="valid")
    for name, r in (("bf16 operands", qrun), ("bf16 / tf32-scoring operands", mix), ("tf32 operands", tf), ("1e-6 perturbed #1", pert[0])):
        dl1 = np.abs(sm(r[:, 2]) - sm(ref[:, 2])) / sm(ref[:, 2])
        dri = np.abs(sm(r[:, 0] + r[:, 1]) - sm(ref[:, 0] + ref[:, 1])) / sm(ref[:, 0] + ref[:, 1]).max()
        print(f"{name}: smoothed L1 curve deviates {dl1.max():.4f} (last {dl1[-1]:.4f}), replace+inject {dri.max():.4f}")
    np.savez_compressed(f"{OUT}/loops_q_c1.npz", q_bf16=qrun.astype(np.float64), q_bf16_tf32=mix.astype(np.float64), q_tf32=tf.astype(np.float64),
                        perturbed_1e6=pert.astype(np.float64))
