"""Phase times (clock64 of CTA 0, frames 0 and 1) of the bf16 whole-step critic kernel (csrc/hg_critic.cu) and its launch time
next to the TF32 kernel's (csrc/critic_fused.cu).  GPU only.
    python tools/critic_trace.py [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cgs_b200.ops as ops, cgs_b200.synth as synth
from cgs_b200 import _lib
from cgs_b200.nets import NewCritic
from cgs_b200.train_handler import FlatAdam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
DEV = "cuda"
ops.set_precision("tf32")
torch.manual_seed(0)
c = NewCritic(dropout=0.3).to(DEV).train()
opt = FlatAdam(c.parameters())
X, Yl, _ = synth.synthetic_frames(B, seed=0)
Xd, Yd = torch.from_numpy(X).to(DEV), torch.from_numpy(Yl[1, :B]).float().to(DEV)
L = _lib.lib()
tr = torch.zeros(64, dtype=torch.int64, device=DEV)


def step(bf16):
    opt.zero_grad()
    ops.critic_train_fused(c, Xd, Yd, 3, rng=c._dropout_rng(DEV), fuse_adam=True, bf16=bf16)
    opt.step()


for it in range(3):
    if it == 2:
        L.cgs_hg_set_trace_critic(ctypes.c_void_p(tr.data_ptr()))
    step(True)
torch.cuda.synchronize()
L.cgs_hg_set_trace_critic(None)
t = tr.cpu().numpy().reshape(2, 32)
names = ["stage + masks", "F0 conv0", "F1 conv1 + clears", "F2 conv2", "F3 conv3", "F4-F6 head", "B5 B4 head", "B3 conv3 w/dgrad",
         "B2 conv2 dgrad || wgrad", "B1 conv1 wgrad || dgrad", "B0 conv0 wgrad"]
print(f"bf16 critic step, B={B}, grid {L.cgs_critic_fused_grid(B)}: clk per phase (frame 0 | frame 1)")
for k, name in enumerate(names):
    print(f"  {name:26s} {t[0, k + 1] - t[0, k]:8d} {t[1, k + 1] - t[1, k]:8d}")
print(f"  {'frame':26s} {t[0, 11] - t[0, 0]:8d} {t[1, 11] - t[1, 0]:8d}   ({(t[1, 11] - t[1, 0]) / 1.965e3:.1f} us)")
for name, bf in (("bf16 (hg_critic.cu)", True), ("tf32 (critic_fused.cu)", False)):
    for _ in range(5):
        step(bf)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(50):
        step(bf)
    e.record(); torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e) / 50 * 1e3:.1f} us per eager step at B={B} (incl. host launch overhead)")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            step(bf)
    g.replay(); torch.cuda.synchronize(); s.record()
    for _ in range(10):
        g.replay()
    e.record(); torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e) / 100 * 1e3:.1f} us per step in a 10-step graph at B={B}")
