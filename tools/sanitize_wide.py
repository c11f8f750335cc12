"""Small run of every wide-path kernel for compute-sanitizer (memcheck / racecheck).  GPU only.
    compute-sanitizer --tool memcheck python tools/sanitize_wide.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgs_b200.ops as ops, cgs_b200.synth as synth
from cgs_b200 import wide
from cgs_b200.nets import NewCritic
from cgs_b200.train_handler import FlatAdam
ops.set_precision("tf32")
torch.manual_seed(0)
DEV = "cuda"
for K, B in ((5, 5), (2, 3)):
    c = NewCritic(chfak=K, dropout=0.3).to(DEV).train()
    X, Y, _ = synth.synthetic_frames(B, seed=K)
    Xd, Yd = torch.from_numpy(X).to(DEV), torch.from_numpy(Y[1, :B]).float().to(DEV)
    loss, pred = wide.critic_train_wide(c, Xd, Yd, 3, rng=c._dropout_rng(DEV))
    torch.cuda.synchronize()
    print("wide chfak", K, "loss", float(loss), "status ok", wide.status_ok())
c = NewCritic(dropout=0.3).to(DEV).train()
opt = FlatAdam(c.parameters())
X, Y, _ = synth.synthetic_frames(7, seed=1)
opt.zero_grad()
l, _ = ops.critic_train_fused(c, torch.from_numpy(X).to(DEV), torch.from_numpy(Y[1, :7]).float().to(DEV), 2, rng=c._dropout_rng(DEV), fuse_adam=True, bf16=True)
opt.step()
torch.cuda.synchronize()
print("bf16 critic loss", float(l))
