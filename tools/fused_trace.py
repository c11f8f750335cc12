"""Phase timing of the whole-step critic kernel (CTA 0, thread 0): cycles per phase for its first frames.
Run on the GPU box:  python tools/fused_trace.py [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgs_b200 import ops, _lib
from cgs_b200.nets import NewCritic

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 592
ops.set_precision("tf32")
L = _lib.lib()
torch.manual_seed(0)
c = NewCritic(dropout=0.3).cuda().train()
from cgs_b200.train_handler import FlatAdam
opt = FlatAdam(c.parameters())          # gradient leaves the kernel as per-CTA partial vectors
X = torch.randint(0, 255, (B, 64, 64, 3), dtype=torch.uint8, device="cuda")
Y = torch.rand(B, device="cuda")
masks = c._dropout_masks(B, X.device)
FULL = "--full" in sys.argv          # as Handler.critic_step launches it: masks drawn in-kernel, Adam in-kernel
if FULL:
    run = lambda: (opt.zero_grad(), ops.critic_train_fused(c, X, Y, 3, rng=c._dropout_rng(X.device), fuse_adam=True), opt.step())
else:
    run = lambda: ops.critic_train_fused(c, X, Y, 3, masks)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if "--flush" in sys.argv else None
for _ in range(3):
    run()
tr = torch.zeros(4 * 24 + 4 * 160, dtype=torch.int64, device="cuda")
if flush is not None:
    flush.zero_(); torch.cuda.synchronize()
L.cgs_critic_fused_set_trace(tr.data_ptr()); run(); torch.cuda.synchronize(); L.cgs_critic_fused_set_trace(None)
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); [run() for _ in range(20)]; t.record(); torch.cuda.synchronize()
print(f"B={B}: {s.elapsed_time(t) / 20 * 1e3:.1f} us/launch")
names = ["F0a stage", "F0 conv0", "F1 conv1", "F2 conv2", "F3 conv3", "F4 4x4", "F5 lin1", "F6 head", "B5 lin1", "B4 4x4",
         "B3 L3", "B2 L2", "B1 L1", "B0a stage", "B0 wgrad0"]
T = tr.cpu()[:96].view(4, 24)
C = tr.cpu()[96:].view(160, 4)
C = C[C[:, 0] != 0]
t0 = int(C[:, 0].min())
print(f"CTAs {len(C)}: start skew {int(C[:,0].max()) - t0} ns; frames done at {int(C[:,1].min()) - t0}..{int(C[:,1].max()) - t0} ns; end {int(C[:,2].min()) - t0}..{int(C[:,2].max()) - t0} ns")
print(f"CTA0: {int(C[0,2]-C[0,0])} ns wall for {int(T[0][16]-T[0][23]) if T[0][16] else 0} clk")
print(f"prologue (launch -> first frame): {int(T[0][0] - T[0][23])} clk")
for f in range(3):
    r = T[f]
    if r[15] == 0:
        break
    print(f"frame {f}: total {int(r[15] - r[0])} clk")
    print("   " + "  ".join(f"{n} {int(r[i + 1] - r[i])}" for i, n in enumerate(names)))
last = max(f for f in range(4) if T[f][15] != 0) if any(T[f][15] != 0 for f in range(4)) else 0
fl = [(int(T[f][17]), int(T[f][16])) for f in range(4) if T[f][16] != 0]
er = [f for f in range(4) if T[f][16] != 0]
if er and T[er[0]][18] != 0:
    E = T[er[0]]
    print(f"in-kernel Adam: partial write + grid barrier {int(E[18] - E[17])} clk, slice sum + update {int(E[19] - E[18])} clk, "
          f"rest {int(E[16] - E[19])} clk")
if fl:
    print(f"end-of-CTA reduce: {fl[0][0] - int(T[last][15])} clk, gradient write: {fl[0][1] - fl[0][0]} clk")
