"""A short program for ncu: a few eager frozen-critic Hourglass steps (batch 1024) and -process inference batches (batch 256)
through the public Handler API.  GPU only.   python tools/prof_hg.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgs_b200.ops as ops, cgs_b200.synth as synth
from cgs_b200.train_handler import FlatAdam, Handler, parse_args
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
DEV = "cuda"
ops.set_precision("tf32")
torch.manual_seed(0)
H = Handler(parse_args(["-frozen"]), device=DEV)
H.critic.to(DEV).train(); H.masker.to(DEV).train()
for q in H.critic.parameters():
    q.requires_grad_(False)
opti = FlatAdam(list(H.masker.parameters()))
X, _, _ = synth.synthetic_frames(2048, seed=0)
Xa, Xb = torch.from_numpy(X[:1024]).to(DEV), torch.from_numpy(X[1024:]).to(DEV)
for it in range(steps):
    t = H.segmentation_step(Xa, Xb, None, opti, roll=it)
H.critic.eval(); H.masker.eval()
for it in range(steps):
    with torch.no_grad():
        H.segment_device(Xa[:256], 0.1)
torch.cuda.synchronize()
print("ok", {k: float(v) for k, v in t.items()})
