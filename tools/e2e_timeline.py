"""GPU timeline of the chunked end-to-end critic pipeline: when do the copies and the step graphs of each chunk start / end?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cgs_b200 import ops
from cgs_b200.graph_step import PipelinedCriticTrainer
from cgs_b200.train_handler import Handler, parse_args
import cgs_b200.synth as synth
B, chunk, nchunks = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 2, 12
ops.set_precision("tf32")
H = Handler(parse_args([]), device="cuda")
X, Y, _ = synth.synthetic_frames(B, seed=0)
Xh = torch.from_numpy(np.concatenate([X] * (2 * chunk))).pin_memory()
Yh = torch.from_numpy(np.tile(Y[1, :B], 2 * chunk)).float().pin_memory()
tr = PipelinedCriticTrainer(H, B)
tr.train(Xh, Yh, chunk=chunk); torch.cuda.synchronize()
main = torch.cuda.current_stream(); cs = tr.copy_stream
ev = lambda: torch.cuda.Event(enable_timing=True)
t0 = ev(); t0.record(main); rows = []; host = []
h0 = time.perf_counter()
for c in range(nchunks):
    sl = tr._cslots[c % 2]
    a, b, g0, g1 = ev(), ev(), ev(), ev()
    hs = time.perf_counter()
    with torch.cuda.stream(cs):
        cs.wait_event(sl["done"]); a.record(cs)
        sl["X"].copy_(Xh[:chunk * B], non_blocking=True); sl["Y"].copy_(Yh[:chunk * B], non_blocking=True)
        b.record(cs); sl["ready"].record(cs)
    main.wait_event(sl["ready"]); g0.record(main); sl["graph"].replay(); g1.record(main); sl["done"].record(main)
    tr.loss_ring[:chunk].copy_(sl["out"], non_blocking=True)
    host.append((time.perf_counter() - hs) * 1e6); rows.append((a, b, g0, g1))
torch.cuda.synchronize()
print(f"chunk={chunk}: host issue time per chunk {np.mean(host):.0f} us; total wall {(time.perf_counter()-h0)*1e3:.2f} ms for {nchunks*chunk} steps")
for c, (a, b, g0, g1) in enumerate(rows):
    print(f"  chunk {c:2d}: copy {t0.elapsed_time(a)*1e3:8.0f} -> {t0.elapsed_time(b)*1e3:8.0f} us | steps {t0.elapsed_time(g0)*1e3:8.0f} -> {t0.elapsed_time(g1)*1e3:8.0f} us")
