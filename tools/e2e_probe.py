"""Where does the end-to-end critic pipeline spend its time?  copies only / graphs only / both.  Run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cgs_b200 import ops
from cgs_b200.graph_step import PipelinedCriticTrainer
from cgs_b200.train_handler import Handler, parse_args
import cgs_b200.synth as synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ops.set_precision("tf32")
H = Handler(parse_args([]), device="cuda")
X, Y, _ = synth.synthetic_frames(B, seed=0)
nb = 2 * chunk
Xh = torch.from_numpy(np.concatenate([X] * nb)).pin_memory()
Yh = torch.from_numpy(np.tile(Y[1, :B], nb)).float().pin_memory()
tr = PipelinedCriticTrainer(H, B)
tr.train(Xh, Yh, chunk=chunk); torch.cuda.synchronize()
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
sl = tr._cslots
t_both = timeit(lambda: tr.train(Xh, Yh, chunk=chunk))
def copies():
    for s in sl:
        s["X"].copy_(Xh[:chunk * B], non_blocking=True); s["Y"].copy_(Yh[:chunk * B], non_blocking=True)
def graphs():
    for s in sl: s["graph"].replay()
t_copy, t_graph = timeit(copies), timeit(graphs)
per = lambda t: t / nb * 1e6
print(f"B={B} chunk={chunk}: per step  both {per(t_both):.1f} us | copies only {per(t_copy):.1f} us ({nb * B * 12292 / t_copy / 1e9:.1f} GB/s) | graphs only {per(t_graph):.1f} us")
