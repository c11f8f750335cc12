"""BASELINE configs[4]: mask-inference throughput sweep, batch 1k-64k synthetic frames resident on the device (uint8 in, fp32 mask
+ uint8 hard mask out), two whole-frame kernels per batch; CUDA events around 5 back-to-back batches.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgs_b200 import ops
from cgs_b200.nets import NewCritic, UnetDecoder
ops.set_precision("tf32")
torch.manual_seed(0)
c, m = NewCritic().cuda().eval(), UnetDecoder().cuda().eval()
print("| frames per batch | ms per batch | M frames/s | GB/s of compulsory traffic (12,288 in + 20,480 out B/frame) |\n|---|---|---|---|")
for B in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
    X = torch.randint(0, 255, (B, 64, 64, 3), dtype=torch.uint8, device="cuda")
    def run():
        pred, o0 = ops.infer_encode_decode(c, m, X)
        return ops.masker_fused(m, X, o0, 0.1)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        run()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"| {B} | {ms:.3f} | {B / ms / 1e3:.2f} | {B * 32768 / ms / 1e6:.0f} |", flush=True)
    del X
