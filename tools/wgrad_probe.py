"""Where does the wgrad kernel's time go?  Times the pipelined wgrad on the three encoder shapes with the final
REDs on/off and different grid sizes (debug env knobs read by the launcher).  Run on the GPU box."""
import os, sys, subprocess
if len(sys.argv) == 1:
    out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, CGS_WGRAD_NOPIPE="1"), capture_output=True, text=True)
    print(f"staged persistent wgrad_mma kernel   : {out.stdout.strip()} {out.stderr.strip()[-200:]}")
    for grid, csz in (("148", "2"),):
        for nored in ("0", "1"):
            env = dict(os.environ, CGS_WGRAD_GRID=grid, CGS_WGRAD_NORED=nored, CGS_WGRAD_CLUSTER=csz)
            out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            print(f"grid {grid:>4s} cluster {csz} nored {nored}: {out.stdout.strip()} {out.stderr.strip()[-200:]}")
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgs_b200 import ops
from cgs_b200._lib import SRC_PLAIN, SRC_POOLBWD
ops.set_precision("tf32")
B = 256
res = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for H, Cin in ((64, 3), (32, 8), (16, 8)):
    x = torch.rand(B, H, H, Cin, device="cuda"); e = torch.rand(B, H // 2, H // 2, 8, device="cuda"); de = torch.rand_like(e)
    idx = torch.randint(0, 4, e.shape, device="cuda", dtype=torch.uint8)
    dw = torch.zeros(8, Cin, 3, 3, device="cuda"); db = torch.zeros(8, device="cuda")
    run = lambda: ops.wgrad3x3(ops._src(SRC_PLAIN, Cin, x), ops._src(SRC_POOLBWD, 8, de, e, idx), B, H, H, dw, db)
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
        s.record(); run(); t.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(t) * 1e3)
    res.append(f"H{H}:{sorted(ts)[len(ts)//2]:6.1f}us")
print("  ".join(res))
