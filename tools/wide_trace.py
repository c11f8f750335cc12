"""clock64() trace of CTA 0 of the TMA / tcgen05 conv kernel (csrc/wide_tc.cu).  GPU only.   python tools/wide_trace.py [B] [hw] [cin] [cout]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgs_b200.ops as ops
from cgs_b200 import wide, _lib
B, hw, cin, cout = (int(a) for a in (sys.argv[1:5] + ["2048", "32", "40", "40"][len(sys.argv) - 1:]))
DEV = "cuda"
ops.set_precision("tf32")
x = wide.to_planar(torch.randn(B, cin, hw, hw, device=DEV)); w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05; b = torch.randn(cout, device=DEV)
tr = torch.zeros(64, dtype=torch.int64, device=DEV)
L = _lib.lib()
for _ in range(2):
    wide.conv3x3(x, w, b, wide.EPI_RELU_POOL)
L.cgs_wide_set_trace(ctypes.c_void_p(tr.data_ptr()))
(pk,) = wide.pack_weights([(w, False)])
wide.conv3x3(x, w, b, wide.EPI_RELU_POOL, packed=pk)
torch.cuda.synchronize()
L.cgs_wide_set_trace(None)
s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        wide.conv3x3(x, w, b, wide.EPI_RELU_POOL, packed=pk)
g.replay(); torch.cuda.synchronize(); s_.record(); g.replay(); e_.record(); torch.cuda.synchronize()
print(f"{s_.elapsed_time(e_) * 100:.1f} us per launch (10 launches in a graph)")
t = tr.cpu().numpy()
t0 = t[56]
print(f"conv {cin}->{cout} {hw}x{hw} B={B}: clk since kernel start, CTA 0")
print(" tile | MMA: top  tmem-free  operands  issued | EPI: wait  ready  stored")
print(f" setup done {t[57] - t0}, MMA thread done {t[58] - t0}, epilogue warp done {t[59] - t0}, CTA done {t[60] - t0}")
for i in range(8):
    r = t[i * 8:i * 8 + 7] - t0
    print(f"  {i:3d} | {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d} {r[6]:8d}")
