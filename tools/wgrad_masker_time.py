"""Time the masker weight-gradient launches of the Hourglass step in isolation (B=1024): masker.0 (cat(x, up(o0)) 11 -> 16,
LeakyReLU grad) and masker.2 (16 -> 1, sigmoid grad), tf32 mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgs_b200 import ops
from cgs_b200._lib import SRC_CATUP, SRC_LEAKYGRAD, SRC_PLAIN, SRC_SIGGRAD
ops.set_precision("tf32")
B = 1024
x = torch.rand(B, 64, 64, 3, device="cuda"); o0 = torch.rand(B, 32, 32, 8, device="cuda")
m0 = torch.rand(B, 64, 64, 16, device="cuda") - 0.3; dm0 = torch.rand_like(m0)
z = torch.rand(B, 64, 64, 1, device="cuda"); dz = torch.rand_like(z)
dw0, db0 = torch.zeros(16, 11, 3, 3, device="cuda"), torch.zeros(16, device="cuda")
dw2, db2 = torch.zeros(1, 16, 3, 3, device="cuda"), torch.zeros(1, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); [fn() for _ in range(n)]; e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
f0 = lambda: ops.wgrad3x3(ops._src(SRC_CATUP, 11, x, o0, C0=3, shift=1), ops._src(SRC_LEAKYGRAD, 16, dm0, m0), B, 64, 64, dw0, db0)
f2 = lambda: ops.wgrad3x3(ops._src(SRC_PLAIN, 16, m0), ops._src(SRC_SIGGRAD, 1, dz, z), B, 64, 64, dw2, db2)
print(f"masker.0 wgrad: {t(f0):.0f} us   masker.2 wgrad: {t(f2):.0f} us   (B={B})")
