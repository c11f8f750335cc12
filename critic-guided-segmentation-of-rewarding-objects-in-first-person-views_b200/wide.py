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


def conv3x3(x, w, bias=None, epi=EPI_PLAIN, transposed=False, idx_in=None, mask=None, want_f32=False):
    """3x3 / padding 1 convolution of a chunk-planar bf16 activation on the tcgen05 kernel.
    transposed=False: w [Cout, Cin, 3, 3] (nets.py:170-183 forward); transposed=True: w is the FORWARD weight [Cx, Cout, 3, 3] of the
    layer whose input gradient this is, x its output gradient with Cx channels.
    Returns out (PLAIN / UNPOOL) or (out, idx[, out_f32 NHWC]) (RELU_POOL)."""
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
            f32 = torch.empty((B, H // 2, W // 2, Cout), device=dev, dtype=torch.float32)
    elif epi == EPI_UNPOOL:
        out = torch.empty((B, Cout // 8, 2 * H, 2 * W, 8), device=dev, dtype=torch.bfloat16)
    else:
        out = torch.empty((B, Cout // 8, H, W, 8), device=dev, dtype=torch.bfloat16)
    _call("cgs_wide_conv3x3", _bp(x), B, H, W, Cin, _p(w.detach()), _p(bias.detach()) if bias is not None else None, Cout,
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
