"""A short program for ncu: a few eager critic training steps at chfak 5 (wide path: TMA / tcgen05 kernels) and at chfak 1 (bf16
whole-step kernel), batch 256, through Handler.critic_step.  GPU only.   python tools/prof_wide.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgs_b200.ops as ops, cgs_b200.synth as synth
from cgs_b200.train_handler import Handler, parse_args
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
DEV = "cuda"
ops.set_precision("tf32")
torch.manual_seed(0)
X, Y, _ = synth.synthetic_frames(256, seed=0)
Xd, Yd = torch.from_numpy(X).to(DEV), torch.from_numpy(Y[1, :256]).float().to(DEV)
for chfak in (5, 1):
    H = Handler(parse_args(["--chfak", str(chfak)]), device=DEV)
    H.critic.to(DEV).train()
    opti = H._opt(H.critic.parameters())
    for it in range(steps):
        l = H.critic_step(Xd, Yd, opti, roll=it)
    torch.cuda.synchronize()
    print("chfak", chfak, "loss", float(l))
