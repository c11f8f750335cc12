import os, sys
sys.path.insert(0, "/root/repo")
import torch
import cgs_b200.ops as ops
from cgs_b200 import wide
ops.set_precision("tf32")
B, C4, K1 = 256, 160, 1280
A = torch.randn(B, K1, device="cuda"); W4 = torch.randn(C4, K1, device="cuda") * 0.03
for _ in range(4):
    wide.gemm(A, True, W4, True, B, C4, K1, relu=True)
torch.cuda.synchronize()
