"""Does H2D copy bandwidth drop while kernels run, and is it specific to the whole-step kernel?  50 MB pinned copies on a side
stream, timed with events, while the main stream runs: nothing / the fused critic step graph / a GEMM loop / a device copy loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgs_b200 import ops
from cgs_b200.graph_step import GraphedCriticStep
from cgs_b200.train_handler import Handler, parse_args
ops.set_precision("tf32")
H = Handler(parse_args([]), device="cuda")
step = GraphedCriticStep(H, 256)
a = torch.rand(4096, 4096, device="cuda"); b = torch.rand(4096, 4096, device="cuda")
big = torch.empty(1 << 28, dtype=torch.uint8, device="cuda"); big2 = torch.empty_like(big)
src = torch.empty(50 << 20, dtype=torch.uint8).pin_memory(); dst = torch.empty(50 << 20, dtype=torch.uint8, device="cuda")
cs = torch.cuda.Stream()
loads = {"idle": lambda: None, "fused critic step": lambda: [step.graph.replay() for _ in range(40)],
         "fp32 GEMM 4096^3": lambda: [torch.mm(a, b) for _ in range(6)], "device copy 256 MB": lambda: [big2.copy_(big) for _ in range(30)]}
for name, fn in loads.items():
    torch.cuda.synchronize()
    fn()                                   # main stream: ~2+ ms of work queued
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        s.record(cs); dst.copy_(src, non_blocking=True); e.record(cs)
    torch.cuda.synchronize()
    print(f"{name:22s}: 50 MB H2D in {s.elapsed_time(e) * 1e3:7.0f} us = {52.4288 / s.elapsed_time(e):5.1f} GB/s")
