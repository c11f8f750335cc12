"""Launch times of the wide-path kernels at two batch sizes (fixed cost vs per-frame slope), CUDA events, 20 launches each.  GPU only.
    python tools/wide_time.py [chfak]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgs_b200.ops as ops
from cgs_b200 import wide
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
DEV = "cuda"
ops.set_precision("tf32")
C = 8 * K


def timeit(fn, n=10):
    """us per launch, n launches replayed from a CUDA graph (no host launch overhead in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(3):
        g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / (3 * n) * 1e3


rows = []
for B in (256, 2048):
    g = torch.Generator(device=DEV).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    r = {}
    for name, cin, cout, hw in (("conv1 fwd 32x32", C, C, 32), ("conv2 fwd 16x16", C, C, 16), ("conv3 fwd 8x8", C, 2 * C, 8)):
        x = wide.to_planar(rnd(B, cin, hw, hw)); w = rnd(cout, cin, 3, 3) * 0.05; b = rnd(cout)
        (pk,) = wide.pack_weights([(w, False)])
        r[name] = timeit(lambda: wide.conv3x3(x, w, b, wide.EPI_RELU_POOL, packed=pk))
    for name, cx, cout, hw in (("conv3 dgrad 8x8 unpool", 2 * C, C, 8), ("conv2 dgrad 16x16 unpool", C, C, 16), ("conv1 dgrad 32x32 plain", C, C, 32)):
        x = wide.to_planar(rnd(B, cx, hw, hw)); w = rnd(cx, cout, 3, 3) * 0.05
        idx = torch.randint(0, 5, (B, cout // 8, hw, hw, 8), device=DEV, dtype=torch.uint8)
        (pk,) = wide.pack_weights([(w, True)])
        if "plain" in name:
            r[name] = timeit(lambda: wide.conv3x3(x, w, transposed=True, packed=pk))
        else:
            r[name] = timeit(lambda: wide.conv3x3(x, w, epi=wide.EPI_UNPOOL, transposed=True, idx_in=idx, packed=pk))
    for name, cin, cout, hw in (("conv1 wgrad 32x32", C, C, 32), ("conv2 wgrad 16x16", C, C, 16), ("conv3 wgrad 8x8", C, 2 * C, 8)):
        x = wide.to_planar(rnd(B, cin, hw, hw)); dy = wide.to_planar(rnd(B, cout, hw, hw))
        dw = torch.zeros(cout, cin, 3, 3, device=DEV); db = torch.zeros(cout, device=DEV)
        r[name] = timeit(lambda: wide.wgrad3x3(x, dy, dw, db))
    X = torch.randint(0, 256, (B, 64, 64, 3), device=DEV, dtype=torch.uint8)
    w0 = rnd(C, 3, 3, 3) * 0.2; b0 = rnd(C)
    r["conv0 fwd"] = timeit(lambda: wide.conv0_fwd(X, 3, w0, b0))
    e0, idx0 = wide.conv0_fwd(X, 3, w0, b0)
    dw0 = torch.zeros_like(w0); db0 = torch.zeros_like(b0)
    r["conv0 wgrad"] = timeit(lambda: wide.conv0_wgrad(X, 3, e0, idx0, dw0, db0))
    C4, K1 = 4 * C, 32 * C
    A = rnd(B, K1); W4 = rnd(C4, K1) * 0.03; H = rnd(B, C4)
    r["gemm H1 = X3 W4^T"] = timeit(lambda: wide.gemm(A, True, W4, True, B, C4, K1, relu=True))
    r["gemm H1, split-K 8"] = timeit(lambda: wide.gemm(A, True, W4, True, B, C4, K1, relu=True, splits=8))
    Wl = rnd(C4, C4) * 0.1
    r["gemm V = H1 Wl1^T"] = timeit(lambda: wide.gemm(H, True, Wl, True, B, C4, C4, relu=True))
    r["gemm dH1 = dV Wl1 (gated)"] = timeit(lambda: wide.gemm(H, True, Wl, False, B, C4, C4, gate=H))
    dWl = torch.zeros(C4, C4, device=DEV)
    r["gemm dWl1 += dV^T H1"] = timeit(lambda: wide.gemm(H, False, H, False, C4, C4, B, out=dWl, accumulate=True))
    (pk,) = wide.pack_weights([(w, True)])
    r["pack (1 job)"] = timeit(lambda: wide.pack_weights([(w, True)]))
    m = ops.dropout_masks([(B, 8, 8, C), (B, 4, 4, 2 * C), (B, 4 * C)], 0.3, 1, torch.zeros(2, dtype=torch.int64, device=DEV))
    st = torch.zeros(2, dtype=torch.int64, device=DEV)
    r["dropout masks"] = timeit(lambda: ops.dropout_masks([(B, 8, 8, C), (B, 4, 4, 2 * C), (B, 4 * C)], 0.3, 1, st))
    r["gemm dE3 = dH1 W4"] = timeit(lambda: wide.gemm(H, True, W4, False, B, K1, C4))
    dW = torch.zeros(C4, K1, device=DEV)
    r["gemm dW4 += dH1^T X3"] = timeit(lambda: wide.gemm(H, False, A, False, C4, K1, B, out=dW, accumulate=True))
    rows.append(r)
print(f"chfak {K}: us per launch at B=256 | B=2048 | per-256-frame slope | fixed")
for k in rows[0]:
    a, b = rows[0][k], rows[1][k]
    slope = (b - a) / 7.0
    print(f"  {k:28s} {a:8.1f} {b:8.1f} {slope:8.1f} {a - slope:8.1f}")
