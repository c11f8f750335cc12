"""Host side of the wide (chfak > 1) path: TMA-fed tcgen05 convolutions on bf16 chunk-planar activations (csrc/wide_tc.cu).

Layout: an activation [B, C, H, W] (C a multiple of 8) is stored as bf16 [B, C/8, H, W, 8]: one pixel's 8 channels of a plane are
one 16-byte slot, which is what both the TMA boxes and the UMMA canonical layouts want (see csrc/wide_tc.cu)."""
import torch

from . import _lib
from .ops import _call, _p, _stream, CgsError

EPI_PLAIN, EPI_RELU_POOL, EPI_UNPOOL = 0, 1, 2


def to_planar(x_nchw):
    """fp32 [B, C, H, W] -> bf16 chunk-planar [B, C/8, H, W, 8]."""
    B, Cc, H, W = x_nchw.shape
    assert Cc % 8 == 0
    return x_nchw.reshape(B, Cc // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def from_planar(t):
    """bf16 / uint8 chunk-planar [B, C/8, H, W, 8] -> [B, C, H, W] (fp32 for bf16 input)."""
    B, CP, H, W, _ = t.shape
    out = t.permute(0, 1, 4, 2, 3).reshape(B, CP * 8, H, W)
    return out.float() if t.dtype == torch.bfloat16 else out.contiguous()


def _bp(t):
    return _p(t, torch.bfloat16)


def pack_weights(jobs):
    """jobs: (w fp32 OIHW, transposed) -> the bf16 operand tiles of cgs_wide_conv3x3 for each, written by ONE launch."""
    import ctypes as C
    L = _lib.lib()
    outs, arr = [], (_lib.WidePackJob * len(jobs))()
    for i, (w, tr) in enumerate(jobs):
        cin, cout = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0])
        o = torch.empty(int(L.cgs_wide_packed_bytes(cin, cout)), device=w.device, dtype=torch.uint8)
        outs.append(o)
        arr[i] = _lib.WidePackJob(_p(w.detach()), _p(o, torch.uint8), cin, cout, int(bool(tr)))
    _call("cgs_wide_pack", C.cast(arr, C.c_void_p), len(jobs), _stream())
    return outs


def conv3x3(x, w, bias=None, epi=EPI_PLAIN, transposed=False, idx_in=None, mask=None, want_f32=False, packed=None):
    """3x3 / padding 1 convolution of a chunk-planar bf16 activation on the tcgen05 kernel.
    transposed=False: w [Cout, Cin, 3, 3] (nets.py:170-183 forward); transposed=True: w is the FORWARD weight [Cx, Cout, 3, 3] of the
    layer whose input gradient this is, x its output gradient with Cx channels.
    Returns out (PLAIN / UNPOOL) or (out, idx[, out_f32 NCHW]) (RELU_POOL)."""
    B, CPi, H, W, _ = x.shape
    Cin = CPi * 8
    Cout = w.shape[1] if transposed else w.shape[0]
    assert (w.shape[0] if transposed else w.shape[1]) == Cin, (w.shape, Cin, transposed)
    dev = x.device
    idx = f32 = None
    if epi == EPI_RELU_POOL:
        out = torch.empty((B, Cout // 8, H // 2, W // 2, 8), device=dev, dtype=torch.bfloat16)
        idx = torch.empty((B, Cout // 8, H // 2, W // 2, 8), device=dev, dtype=torch.uint8)
        if want_f32:
            f32 = torch.empty((B, Cout, H // 2, W // 2), device=dev, dtype=torch.float32)
    elif epi == EPI_UNPOOL:
        out = torch.empty((B, Cout // 8, 2 * H, 2 * W, 8), device=dev, dtype=torch.bfloat16)
    else:
        out = torch.empty((B, Cout // 8, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("cgs_wide_conv3x3", _bp(x), B, H, W, Cin, _p(w.detach()), _p(packed, torch.uint8), _p(bias.detach()) if bias is not None else None, Cout,
          int(bool(transposed)), int(epi), _bp(out), _p(f32), _p(idx, torch.uint8), _p(idx_in, torch.uint8), _p(mask), _stream())
    if epi == EPI_RELU_POOL:
        return (out, idx, f32) if want_f32 else (out, idx)
    return out


_ws = {}


def _workspace(n, dev):
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    t = _ws.get(key)
    if t is None or t.numel() < n:
        t = torch.empty(n, device=dev, dtype=torch.float32)
        _ws[key] = t
    return t


def wgrad3x3(x, dy, dw, db=None):
    """dw [Cout, Cin, 3, 3] += weight gradient, db [Cout] += bias gradient of a 3x3 / padding 1 convolution with input x and output
    gradient dy (both chunk-planar bf16), on the tcgen05 kernel (GEMM with K = pixels)."""
    B, CPi, H, W, _ = x.shape
    Cout = dy.shape[1] * 8
    assert tuple(dy.shape) == (B, Cout // 8, H, W, 8) and tuple(dw.shape) == (Cout, CPi * 8, 3, 3)
    n = int(_lib.lib().cgs_wide_wgrad_workspace(B, H, W, Cout))
    ws = _workspace(n, x.device)
    _call("cgs_wide_wgrad3x3", _bp(x), _bp(dy), B, H, W, CPi * 8, Cout, _p(dw), _p(db), _p(ws), ws.numel(), _stream())


def status_ok():
    return _lib.lib().cgs_wide_status() == 0


# ------------------------------------------------------------------------------------------------------------------------
# the rest of the wide critic step (csrc/wide_misc.cu) and its orchestration
def conv0_fwd(frames_u8, roll, w0, b0):
    """features.0 + ReLU + MaxPool on raw uint8 frames [B,64,64,3] (nets.py:170-172 after main.py:185-189's cast / shift):
    returns (e0 [B, C0/8, 32, 32, 8] bf16, idx0 uint8)."""
    B, C0, dev = frames_u8.shape[0], w0.shape[0], frames_u8.device
    e0 = torch.empty((B, C0 // 8, 32, 32, 8), device=dev, dtype=torch.bfloat16)
    idx0 = torch.empty((B, C0 // 8, 32, 32, 8), device=dev, dtype=torch.uint8)
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    _call("cgs_wide_conv0_fwd", _p(frames_u8, torch.uint8), B, r, rd, _p(w0.detach()), _p(b0.detach()), C0, _bp(e0), _p(idx0, torch.uint8), _stream())
    return e0, idx0


def conv0_wgrad(frames_u8, roll, de0, idx0, dw0, db0):
    """dw0 [C0,3,3,3] +=, db0 [C0] += from the pooled gradient de0 [B, C0/8, 32, 32, 8] and the arg-max bytes."""
    B, C0 = frames_u8.shape[0], dw0.shape[0]
    ws = _workspace(160 * 48 * C0, frames_u8.device)
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    _call("cgs_wide_conv0_wgrad", _p(frames_u8, torch.uint8), B, r, rd, _bp(de0), _p(idx0, torch.uint8), C0, _p(dw0), _p(db0), _p(ws), ws.numel(),
          _stream())


_gemm_ws = {}


def gemm(A, a_kc, Bm, b_kc, M, N, K, out=None, bias=None, gate=None, relu=False, accumulate=False, splits=1):
    """out[M, N] (+)= op(A) @ op(B) in TF32: A is [M, K] (a_kc) or [K, M]; B is [N, K] (b_kc) or [K, N], all row-major fp32.
    splits > 1 cuts K over that many CTAs per output tile (deterministic: the last CTA sums the partial tiles in order)."""
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    assert A.dim() == 2 and Bm.dim() == 2 and tuple(A.shape) == ((M, K) if a_kc else (K, M)) and tuple(Bm.shape) == ((N, K) if b_kc else (K, N))
    ws = ctr = None
    if splits > 1:
        key = A.device.index if A.device.index is not None else torch.cuda.current_device()
        ent = _gemm_ws.get(key)
        if ent is None or ent[0].numel() < splits * M * N:
            ent = (torch.empty(max(splits * M * N, 1 << 20), device=A.device, dtype=torch.float32),
                   torch.zeros(4096, device=A.device, dtype=torch.int32))
            _gemm_ws[key] = ent
        ws, ctr = ent
    _call("cgs_wide_gemm", _p(A), int(a_kc), A.shape[1], _p(Bm), int(b_kc), Bm.shape[1], _p(out), N, M, N, K, _p(bias), _p(gate), int(relu),
          int(accumulate), int(splits), _p(ws), _p(ctr, torch.int32), _stream())
    return out


def colsums(jobs, B):
    """jobs: (X [B, n], out [n], scale, accumulate) x <= 5 -> out = (accumulate ? out : 0) + scale * X.sum(0), one launch."""
    import ctypes as C
    arr = (_lib.WideColJob * len(jobs))(*[_lib.WideColJob(_p(X), _p(o), int(o.numel()), float(sc), int(acc)) for X, o, sc, acc in jobs])
    _call("cgs_wide_colsums", C.cast(arr, C.c_void_p), len(jobs), B, _stream())


def supported(critic):
    """The wide path covers NewCritic at chfak 2..5 (channel counts 8k, 8k, 8k, 16k, 32k with k <= 5) in the tensor-core mode."""
    from . import ops
    f = critic.features
    c = [f[0].out_channels, f[3].out_channels, f[6].out_channels, f[10].out_channels, f[14].out_channels]
    k = c[0] // 8
    return bool(ops._precision) and 2 <= k <= 5 and c == [8 * k, 8 * k, 8 * k, 16 * k, 32 * k] and f[0].in_channels == 3


def critic_train_wide(critic, frames_u8, target, roll=0, masks=(None, None, None), loss_grad=1.0, bce=False, rng=None):
    """Forward + loss + backward of one critic_pipe step (reference main.py:185-198) for a wide NewCritic: features.0 and its
    gradient on bf16 mma.sync, features.3 / .6 / .10 forward, input and weight gradients on TMA-fed tcgen05, the head as TF32 GEMMs.
    The parameter gradients are ADDED to the parameters' `.grad` (fixed summation order: bit-reproducible).
    masks: (m2 [B,8,8,C2], m3 [B,4,4,C3], mv [B,nb]) NHWC fp32 or Nones; rng = (p, seed, state): drawn from the module's Philox
    stream by cgs_dropout_masks instead.  Returns (loss, pred [B])."""
    from . import ops
    params = list(critic.parameters())
    assert len(params) == 14
    w0, b0, w1, b1, w2, b2, w3, b3, w4, b4, wl1, bl1, wl2, bl2 = [q.detach() for q in params]
    g = []
    for q in params:
        if q.grad is None:
            q.grad = torch.zeros_like(q)
        o = getattr(q, "_cgs_opt", None)
        if o is not None and q.grad is getattr(q, "_cgs_grad", None):
            o._clean = False
        g.append(q.grad)
    dw0, db0, dw1, db1, dw2, db2, dw3, db3, dw4, db4, dwl1, dbl1, dwl2, dbl2 = g
    B, dev = frames_u8.shape[0], frames_u8.device
    C2, C3, C4 = w2.shape[0], w3.shape[0], w4.shape[0]
    K1 = 16 * C3
    if rng is not None:
        masks = ops.dropout_masks([(B, 8, 8, C2), (B, 4, 4, C3), (B, C4)], float(rng[0]), rng[1], rng[2])
    m2, m3, mv = masks
    # ---- forward
    p1, p2, p3, p3t, p2t, p1t = pack_weights([(w1, False), (w2, False), (w3, False), (w3, True), (w2, True), (w1, True)])
    e0, idx0 = conv0_fwd(frames_u8, roll, w0, b0)
    e1, idx1 = conv3x3(e0, w1, b1, EPI_RELU_POOL, packed=p1)
    e2, idx2 = conv3x3(e1, w2, b2, EPI_RELU_POOL, mask=m2, packed=p2)
    e3, idx3, e3f = conv3x3(e2, w3, b3, EPI_RELU_POOL, mask=m3, want_f32=True, packed=p3)
    X3 = e3f.view(B, K1)
    W4 = w4.reshape(C4, K1)
    H1 = gemm(X3, True, W4, True, B, C4, K1, bias=b4, relu=True, splits=8)                    # features.14 + ReLU (nets.py:186-187)
    V = gemm(H1, True, wl1, True, B, C4, C4, bias=bl1, relu=True)                      # crit.1 + ReLU (nets.py:190-191)
    pred = torch.empty(B, device=dev, dtype=torch.float32)
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    scr = torch.empty((2 * B * C4 + 2 * B,), device=dev, dtype=torch.float32)
    dV, U, dz, lterm = scr[:B * C4].view(B, C4), scr[B * C4:2 * B * C4].view(B, C4), scr[2 * B * C4:2 * B * C4 + B], scr[2 * B * C4 + B:]
    _call("cgs_wide_head_mid", _p(V), _p(mv), _p(wl2.reshape(-1)), _p(bl2), _p(target), B, C4, float(loss_grad), int(bool(bce)), _p(pred), _p(lterm),
          _p(dV), _p(dz), _p(U), _stream())
    # ---- backward: head
    dH1 = gemm(dV, True, wl1, False, B, C4, C4, gate=H1)                               # through crit.1 and features.15's ReLU
    gemm(dV, False, H1, False, C4, C4, B, out=dwl1, accumulate=True)
    colsums([(dV, dbl1, 1.0, 1), (dH1, db4, 1.0, 1), (U, dwl2.view(-1), 1.0, 1), (dz.view(B, 1), dbl2, 1.0, 1), (lterm.view(B, 1), loss, 1.0 / B, 0)], B)
    dE3 = gemm(dH1, True, W4, False, B, K1, C4)                                        # [B, C3, 4, 4]
    gemm(dH1, False, X3, False, C4, K1, B, out=dw4.view(C4, K1), accumulate=True)
    dY3 = torch.empty((B, C3 // 8, 8, 8, 8), device=dev, dtype=torch.bfloat16)
    _call("cgs_wide_unpool3", _p(dE3), _p(idx3, torch.uint8), _p(m3), B, C3, _bp(dY3), _stream())
    # ---- backward: the 3x3 convolutions
    wgrad3x3(e2, dY3, dw3, db3)
    dY2 = conv3x3(dY3, w3, epi=EPI_UNPOOL, transposed=True, idx_in=idx2, mask=m2, packed=p3t)
    wgrad3x3(e1, dY2, dw2, db2)
    dY1 = conv3x3(dY2, w2, epi=EPI_UNPOOL, transposed=True, idx_in=idx1, packed=p2t)
    wgrad3x3(e0, dY1, dw1, db1)
    dE0 = conv3x3(dY1, w1, transposed=True, packed=p1t)
    conv0_wgrad(frames_u8, roll, dE0, idx0, dw0, db0)
    return loss.reshape(()), pred
