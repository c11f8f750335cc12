"""Phase times (clock64 of CTA 0, frames 0 and 1) of the whole-frame Hourglass kernels.  GPU only.
    python tools/hg_trace.py [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cgs_b200.ops as ops, cgs_b200.synth as synth
from cgs_b200 import _lib
from cgs_b200.nets import NewCritic, UnetDecoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
DEV = "cuda"
ops.set_precision("tf32")
torch.manual_seed(0)
c, m = NewCritic(dropout=0.3).to(DEV).train(), UnetDecoder().to(DEV).train()
X, _, _ = synth.synthetic_frames(B, seed=0)
Xd = torch.from_numpy(X).to(DEV)
tf = torch.zeros(64, dtype=torch.int64, device=DEV); tb = torch.zeros(64, dtype=torch.int64, device=DEV)
ts = torch.zeros(64, dtype=torch.int64, device=DEV)
XB, _, _ = synth.synthetic_frames(B, seed=1)
XBd = torch.from_numpy(XB).to(DEV)
L = _lib.lib()
tape = ops.hg_tape(B, DEV); pack = ops.hg_pack(c, m)
dz = torch.randn(B, 64, 64, device=DEV) / B
for it in range(3):
    if it == 2:
        L.cgs_hg_set_trace(ctypes.c_void_p(tf.data_ptr()), ctypes.c_void_p(tb.data_ptr()), ctypes.c_void_p(ts.data_ptr()))
    pred, z, _ = ops.hg_forward(c, m, Xd, train=True, rng=c._dropout_rng(DEV), tape=tape, pack=pack)
    ops.hg_score_bf16(c, Xd, XBd, z, pack, None, pred.squeeze(1), rng=c._dropout_rng(DEV), l1=0.5)
    ops.hg_backward(m, Xd, tape, z, dz, pack=pack)
torch.cuda.synchronize()
L.cgs_hg_set_trace(None, None, None)
tf, tb = tf.cpu().numpy().reshape(2, 32), tb.cpu().numpy().reshape(2, 32)
fn = ["stage", "conv0", "conv1", "conv2", "conv3", "head+dec4", "dec3", "dec2", "dec1", "dec0", "tape", "band prep",
      "band0 masker.0", "band0 P (MMA)", "band0 stencil"]
print(f"forward, B={B}: clk per phase (frame 0 | frame 1)")
for k, name in enumerate(fn):
    print(f"  {name:14s} {tf[0, k + 1] - tf[0, k]:8d} {tf[1, k + 1] - tf[1, k]:8d}")
print(f"  {'bands 1-7':14s} {tf[0, 20] - tf[0, 15]:8d} {tf[1, 20] - tf[1, 15]:8d}")
print(f"  {'frame':12s} {tf[0, 20] - tf[0, 0]:8d} {tf[1, 20] - tf[1, 0]:8d}   ({(tf[1, 20] - tf[1, 0]) / 1.965e3:.1f} us)")
bn = ["tape load", "-"] + [f"band{b} {w}" for b in range(4) for w in ("B1 stage", "B2 m0", "B3 m2 wgrad", "B4 m2 dgrad", "B5 m0 w/dgrad")] + \
     ["D0 dec0", "D1 dec1", "D2 dec2", "D3 dec3", "D4 dec4"]
print(f"backward, B={B}: clk per phase (frame 0 | frame 1)")
for k, name in enumerate(bn):
    print(f"  {name:18s} {tb[0, k + 1] - tb[0, k]:8d} {tb[1, k + 1] - tb[1, k]:8d}")
print(f"  {'frame':18s} {tb[0, 27] - tb[0, 0]:8d} {tb[1, 27] - tb[1, 0]:8d}   ({(tb[1, 27] - tb[1, 0]) / 1.965e3:.1f} us)")
ts = ts.cpu().numpy().reshape(2, 2, 16)
sn = ["stage blend", "F0 conv0", "F1 conv1+zero", "F2 conv2", "F3 conv3", "F4-F6 head", "B5 B4 head", "B3 conv3 dgrad", "B2 conv2 dgrad",
      "B1 conv1 dgrad", "B0 conv0 dgrad"]
print(f"scoring, B={B}: clk per phase, frame 1 (replace pass | inject pass)")
for k, name in enumerate(sn):
    print(f"  {name:18s} {ts[1, 0, k + 1] - ts[1, 0, k]:8d} {ts[1, 1, k + 1] - ts[1, 1, k]:8d}")
print(f"  {'pass':18s} {ts[1, 0, 11] - ts[1, 0, 0]:8d} {ts[1, 1, 11] - ts[1, 1, 0]:8d}   frame incl. critic(B): {(ts[1, 1, 11] - ts[0, 1, 11]) / 1.965e3:.1f} us")
for name, fnc in (("scoring (critic(B) + 2 blends)", lambda: ops.hg_score_bf16(c, Xd, XBd, z, pack, None, pred.squeeze(1), rng=c._dropout_rng(DEV), l1=0.5)),
                  ("forward", lambda: ops.hg_forward(c, m, Xd, train=True, rng=c._dropout_rng(DEV), tape=tape, pack=pack)),
                  ("forward (eval, no tape)", lambda: ops.hg_forward(c.eval(), m, Xd, thresh=0.1, pack=pack)),
                  ("backward", lambda: ops.hg_backward(m, Xd, tape, z, dz, pack=pack))):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fnc(); torch.cuda.synchronize(); s.record()
    for _ in range(20):
        fnc()
    e.record(); torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e) / 20 * 1e3:.1f} us per launch at B={B}")
    c.train()
