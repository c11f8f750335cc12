"""Phase timing of the tcgen05 conv kernel (CTA 0, thread 0): cycles spent per tile in
stage(next) | wait MMA | fence+sync | issue(next) | epilogue.   Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgs_b200 import ops, _lib
from cgs_b200._lib import SRC_PLAIN, EPI_RELU_POOL, EPI_LINEAR
ops.set_precision("tf32")
L = _lib.lib()
for (B, H, Cin, Cout, epi) in ((256, 64, 3, 8, EPI_RELU_POOL), (256, 32, 8, 8, EPI_RELU_POOL), (256, 64, 16, 16, EPI_LINEAR)):
    x = torch.rand(B, H, H, Cin, device="cuda"); w = torch.rand(Cout, Cin, 3, 3, device="cuda") - 0.5; b = torch.zeros(Cout, device="cuda")
    pooled = epi == EPI_RELU_POOL
    e = torch.empty(B, H // 2 if pooled else H, H // 2 if pooled else H, Cout, device="cuda")
    idx = torch.empty(e.shape, device="cuda", dtype=torch.uint8) if pooled else None
    tr = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
    run = lambda: ops.conv3x3(ops._src(SRC_PLAIN, Cin, x), w, b, B, H, H, Cout, epi, e, idx_out=idx)
    for _ in range(3): run()
    L.cgs_tc_set_trace(tr.data_ptr()); run(); torch.cuda.synchronize(); L.cgs_tc_set_trace(None)
    s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
    s.record(); [run() for _ in range(20)]; t.record(); torch.cuda.synchronize()
    t_ = tr.cpu().view(16, 8)
    print(f"--- B{B} H{H} {Cin}->{Cout} epi{epi}: {s.elapsed_time(t)/20*1e3:.1f} us/launch")
    for i in range(8):
        r = t_[i]
        if r[5] == 0: break
        print(f" tile{i}: stage {r[1]-r[0]:6d} wait {r[2]-r[1]:6d} sync {r[3]-r[2]:6d} issue {r[4]-r[3]:6d} epi {r[5]-r[4]:6d}  total {r[5]-r[0]:6d}")
