"""Diagnostic (GPU): which dropout masks does pass 1 (injected) of cgs_hg_score use when it draws them in the kernel?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cgs_b200.ops as ops, cgs_b200.synth as synth
from cgs_b200.nets import NewCritic
DEV = "cuda"
ops.set_precision("tf32")
B, p = 40, 0.3
csd = synth.perturbed_state(synth.critic_shapes(1), 71, 1.5)
XA, _, _ = synth.synthetic_frames(B, seed=71); XB, _, _ = synth.synthetic_frames(B, seed=72)
g = torch.Generator().manual_seed(1)
Z = (torch.rand(B, 64, 64, generator=g) * 0.9 + 0.05).to(DEV)
tr, ti = torch.rand(B, generator=g).to(DEV), torch.rand(B, generator=g).to(DEV)
Ad, Bd = torch.from_numpy(XA).to(DEV), torch.from_numpy(XB).to(DEV)
def critic():
    c = NewCritic(dropout=p); c.load_state_dict({k: torch.from_numpy(v) for k, v in csd.items()}); return c.to(DEV).train()
torch.manual_seed(5)
c = critic()
ms = [[t.clone() for t in c._dropout_masks(B, DEV)] for _ in range(4)]
torch.manual_seed(5)
c2 = critic(); c2._instance = c._instance
_, dz2, pr2, pi2 = ops.hg_score(c2, Ad, Bd, Z, tr, ti, roll=3, rng=c2._dropout_rng(DEV), l1=0.5)
print("state after rng call:", c2._rng_state.tolist())
for i in range(4):
    for j in range(4):
        _, dz1, pr1, pi1 = ops.hg_score(c, Ad, Bd, Z, tr, ti, roll=3, masks=ms[i], masks_inject=ms[j], l1=0.5)
        print(f"forced masks (call {i}, call {j}): pred_replace equal {torch.equal(pr1, pr2)}  pred_inject equal {torch.equal(pi1, pi2)}  "
              f"max|dpi| {(pi1 - pi2).abs().max().item():.3e}")
# same kernel, replace-only
_, _, pr3, _ = ops.hg_score(c, Ad, Bd, Z, tr, None, roll=3, masks=ms[0], l1=0.5)
print("replace-only forced == rng pass 0:", torch.equal(pr3, pr2))
