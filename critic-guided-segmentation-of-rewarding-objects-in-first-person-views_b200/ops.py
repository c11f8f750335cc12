"""torch.autograd bindings of the C-ABI kernels (include/cgs_b200.h).

Every tensor handled here is NHWC-contiguous fp32 on a CUDA device ([B,H,W,C]); the
modules in nets.py convert at the boundary (zero-copy for the channels_last inputs the
reference loops produce, main.py:189,360).  torch is used for memory, streams and
autograd bookkeeping only; all arithmetic runs in libcgs_b200.so.  No CPU path.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_LEAKY, EPI_LINEAR, EPI_MUL, EPI_RELU_POOL, EPI_SIGMOID, EPI_SPLIT_UP, SRC_CATUP,
                   SRC_LEAKYGRAD, SRC_PLAIN, SRC_POOLBWD, SRC_SIGGRAD, SRC_U8ROLL, CgsError, Conv3x3Args, Src, Wgrad3x3Args)

_launches = 0   # kernels launched through this module (bench.py reports it)
_weights_epoch = 0   # bumped by every optimizer kernel launched through this module: those write parameters behind torch's
                     # back (no `_version` change), so caches derived from weights (packed decoder fragments) key on it too
_precision = 0  # 0 = fp32 FFMA kernels (exact parity path), 1 = tcgen05 TF32 convolutions where covered


def set_precision(name):
    """'fp32' (default; rtol 1e-4 vs the reference) or 'tf32' (tcgen05 tensor-core convolutions, fp32 accumulate)."""
    global _precision
    _precision = {"fp32": 0, "tf32": 1}[name]


def get_precision():
    return "tf32" if _precision else "fp32"


def weights_changed():
    """Tell the weight-derived caches that parameters were rewritten outside this module's entry points (CUDA-graph replays
    of training steps run the optimizer kernels without passing through here)."""
    global _weights_epoch
    _weights_epoch += 1


def launch_count():
    return _launches


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise CgsError("cgs_b200 kernels need CUDA tensors: there is no CPU fallback")
    if t.dtype != dtype or not t.is_contiguous():
        raise CgsError(f"expected contiguous {dtype} tensor, got {t.dtype} strides {t.stride()}")
    return t.data_ptr()


_prof = None    # list of (entry point, start event, end event) while a profile_calls() block is active


class profile_calls:
    """Times every C-ABI launch issued inside the block with a CUDA-event pair on the launching stream (eager execution
    only; bench.py uses it to find the dominant kernel of a step and its live duration)."""

    def __enter__(self):
        global _prof
        _prof = self.records = []
        return self

    def __exit__(self, *exc):
        global _prof
        _prof = None
        return False

    def summary(self):
        """{entry point: (launches, total milliseconds)} (synchronises)."""
        torch.cuda.synchronize()
        out = {}
        for name, s, e in self.records:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + s.elapsed_time(e))
        return out


def _call(name, *args):
    global _launches
    if _prof is not None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    rc = getattr(_lib.lib(), name)(*args)
    _lib.check(rc, name)
    _launches += 1
    if _prof is not None:
        e.record()
        _prof.append((name, s, e))


def _src(mode, Cn, a, b=None, idx=None, C0=0, shift=0):
    return Src(mode, Cn, C0, shift, _p(a), _p(b), _p(idx, torch.uint8))


def conv3x3(src, w, bias, B, H, W, Cout, epi, out, transposed=False, out2=None, idx_out=None, mul=None,
            C0=0, shift2=0, thresh=0.0):
    a = Conv3x3Args(src, _p(w), _p(bias), int(transposed), B, H, W, Cout, epi, _p(out), _p(out2),
                    _p(idx_out, torch.uint8), _p(mul), C0, shift2, thresh, _precision)
    _call("cgs_conv3x3", C.byref(a), _stream())


def wgrad3x3(xsrc, dysrc, B, H, W, dw, db):
    a = Wgrad3x3Args(xsrc, dysrc, B, H, W, _p(dw), _p(db), _precision)
    _call("cgs_wgrad3x3", C.byref(a), _stream())


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _frames_src(x, roll):
    """Operand descriptor for an encoder input: fp32 NHWC, or uint8 NHWC frames with the /255 cast and the
    shift_batch roll fused into the load (roll: int, or int32 device scalar for graph replay)."""
    if x.dtype == torch.uint8:
        if torch.is_tensor(roll):
            return Src(SRC_U8ROLL, x.shape[3], 0, 0, _p(x, torch.uint8), _p(roll, torch.int32), None)
        return Src(SRC_U8ROLL, x.shape[3], 0, int(roll or 0), _p(x, torch.uint8), None, None)
    return None


# ---- weight gradients on a side stream.  A layer's wgrad and dgrad are independent; each alone leaves most SMs
# idle at the reference batch sizes, so wgrad kernels that accumulate into a FlatAdam bucket are forked onto a second
# stream and joined in FlatAdam.step() (captured as parallel branches of the step graph).
async_wgrad = True
_side = {"stream": None, "pending": False, "keep": []}


def _fork_wgrad(launch, keep, direct):
    """Run `launch()` (wgrad kernels) on the side stream if its outputs go straight into the flat bucket."""
    if not (async_wgrad and direct):
        launch()
        return
    if _side["stream"] is None:
        _side["stream"] = torch.cuda.Stream()
    main, side = torch.cuda.current_stream(), _side["stream"]
    side.wait_stream(main)
    with torch.cuda.stream(side):
        launch()
    _side["pending"] = True
    _side["keep"].extend(t for t in keep if t is not None)    # keep operands alive until the join


def join_wgrad():
    """Make the current stream wait for all forked wgrad kernels (called by FlatAdam.step / zero_grad)."""
    if _side["pending"]:
        torch.cuda.current_stream().wait_stream(_side["stream"])
        _side["pending"] = False
        _side["keep"].clear()


def _gbuf(param, needed):
    """Where a parameter gradient goes: (buffer to ACCUMULATE into, value to return to autograd).
    Parameters owned by a FlatAdam carry `_cgs_grad`, a view of the flat gradient bucket: the kernels add
    into it directly (no zero-fill, no AccumulateGrad add) and autograd gets None."""
    if not needed:
        return None, None
    g = getattr(param, "_cgs_grad", None)
    if g is not None and param.grad is g:      # still the live .grad (not reset by a foreign zero_grad)
        opt = getattr(param, "_cgs_opt", None)
        if opt is not None:
            opt._clean = False                 # the bucket now holds gradient: zero_grad must really clear it
        return g, None
    z = torch.zeros_like(param)
    return z, z


class EncBlock(torch.autograd.Function):
    """[Dropout ->] Conv2d(3,1,1) -> ReLU -> MaxPool2d(2): one NewCritic.features stage
    (reference nets.py:170-183).  x [B,H,W,Cin]; mask = multiplicative dropout mask or None."""

    @staticmethod
    def forward(ctx, x, mask, w, b, roll=None):
        B, H, W, Cin = x.shape
        Cout = w.shape[0]
        e = torch.empty((B, H // 2, W // 2, Cout), device=x.device, dtype=torch.float32)
        idx = torch.empty((B, H // 2, W // 2, Cout), device=x.device, dtype=torch.uint8)
        if x.dtype == torch.uint8 and Cin == 3 and H % 16 == 0 and W % 16 == 0 and Cout % 8 == 0:
            # dedicated first-layer kernel: raw frames in, cast + roll + conv + ReLU + pool fused
            rdev = roll if torch.is_tensor(roll) else None
            _call("cgs_conv_rgb_fwd", _p(x, torch.uint8), B, H, W, 0 if rdev is not None else int(roll or 0),
                  _p(rdev, torch.int32), _p(w), _p(b), Cout, _p(e), _p(idx, torch.uint8), _stream())
        else:
            xs = _frames_src(x, roll) or _src(SRC_PLAIN, Cin, x, mask)
            conv3x3(xs, w, b, B, H, W, Cout, EPI_RELU_POOL, e, idx_out=idx)
        ctx.save_for_backward(x, mask, w, e, idx)
        ctx.roll, ctx.params = roll, (w, b)
        ctx.set_materialize_grads(False)
        return e

    @staticmethod
    def backward(ctx, de):
        x, mask, w, e, idx = ctx.saved_tensors
        if de is None:
            return None, None, None, None, None
        de = _c(de)
        B, H, W, Cin = x.shape
        Cout = w.shape[0]
        dy = _src(SRC_POOLBWD, Cout, de, e, idx)
        dx = rw = rb = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dw, rw = _gbuf(ctx.params[0], True)
            db, rb = _gbuf(ctx.params[1], True)
            xs = _frames_src(x, ctx.roll) or _src(SRC_PLAIN, Cin, x, mask)
            _fork_wgrad(lambda: wgrad3x3(xs, dy, B, H, W, dw, db), (x, mask, de, e, idx, ctx.roll if torch.is_tensor(ctx.roll) else None),
                        rw is None and rb is None)
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            conv3x3(dy, w, None, B, H, W, Cin, EPI_MUL if mask is not None else EPI_LINEAR, dx,
                    transposed=True, mul=mask)
        return dx, None, rw, rb, None


class Head(torch.autograd.Function):
    """NewCritic tail: [Dropout] Conv2d(16c,32c,4) ReLU | Flatten Linear ReLU [Dropout] Linear Sigmoid
    (reference nets.py:183-195).  Returns (pred [B,1], e4 [B,1,1,NB])."""

    @staticmethod
    def forward(ctx, e3, m_e3, m_v, w14, b14, w1, b1, w2, b2):
        B, _, _, C3 = e3.shape
        NB = w14.shape[0]
        e4 = torch.empty((B, 1, 1, NB), device=e3.device, dtype=torch.float32)
        v = torch.empty((B, NB), device=e3.device, dtype=torch.float32)
        pred = torch.empty((B, 1), device=e3.device, dtype=torch.float32)
        _call("cgs_head_fwd", _p(e3), _p(m_e3), _p(m_v), _p(w14), _p(b14), _p(w1), _p(b1), _p(w2), _p(b2),
              B, C3, NB, _p(e4), _p(v), _p(pred), _stream())
        ctx.save_for_backward(e3, m_e3, m_v, w14, w1, w2, e4, v, pred)
        ctx.params = (w14, b14, w1, b1, w2, b2)
        ctx.set_materialize_grads(False)
        return pred, e4

    @staticmethod
    def backward(ctx, dpred, de4):
        e3, m_e3, m_v, w14, w1, w2, e4, v, pred = ctx.saved_tensors
        B, _, _, C3 = e3.shape
        NB = w14.shape[0]
        if dpred is None and de4 is None:
            return (None,) * 9
        dpred = torch.zeros_like(pred) if dpred is None else _c(dpred)
        de4 = None if de4 is None else _c(de4)
        want_w = any(ctx.needs_input_grad[3:])
        bufs, rets = zip(*[_gbuf(prm, want_w) for prm in ctx.params])
        de3 = torch.empty_like(e3) if ctx.needs_input_grad[0] else None
        _call("cgs_head_bwd", _p(e3), _p(m_e3), _p(m_v), _p(w14), _p(w1), _p(w2), _p(e4), _p(v), _p(pred),
              _p(dpred), _p(de4), B, C3, NB, *[_p(t) for t in bufs], _p(de3), _stream())
        return (de3, None, None) + tuple(rets)


def tail_supported(B, C2, C3, NB):
    return bool(_lib.lib().cgs_tail_supported(B, C2, C3, NB))


class Tail(torch.autograd.Function):
    """features[9..15] + crit of NewCritic (reference nets.py:179-195) fused: [Dropout] Conv3x3 ReLU MaxPool (-> embeds[3])
    [Dropout] Conv4x4 ReLU (-> embeds[4]) Flatten Linear ReLU [Dropout] Linear Sigmoid.  Returns (pred, e3, e4)."""

    @staticmethod
    def forward(ctx, e2, m_e2, m_e3, m_v, w3, b3, w14, b14, w1, b1, w2, b2):
        B, _, _, C2 = e2.shape
        C3, NB = w3.shape[0], w14.shape[0]
        dev = e2.device
        e3 = torch.empty((B, 4, 4, C3), device=dev, dtype=torch.float32)
        idx3 = torch.empty((B, 4, 4, C3), device=dev, dtype=torch.uint8)
        e4 = torch.empty((B, 1, 1, NB), device=dev, dtype=torch.float32)
        v = torch.empty((B, NB), device=dev, dtype=torch.float32)
        pred = torch.empty((B, 1), device=dev, dtype=torch.float32)
        _call("cgs_tail_fwd", _p(e2), _p(m_e2), _p(m_e3), _p(m_v), _p(w3), _p(b3), _p(w14), _p(b14), _p(w1), _p(b1),
              _p(w2), _p(b2), B, C2, C3, NB, _p(e3), _p(idx3, torch.uint8), _p(e4), _p(v), _p(pred), _stream())
        ctx.save_for_backward(e2, m_e2, m_e3, m_v, w3, w14, w1, w2, e3, idx3, e4, v, pred)
        ctx.params = (w3, b3, w14, b14, w1, b1, w2, b2)
        ctx.set_materialize_grads(False)
        return pred, e3, e4

    @staticmethod
    def backward(ctx, dpred, de3, de4):
        e2, m_e2, m_e3, m_v, w3, w14, w1, w2, e3, idx3, e4, v, pred = ctx.saved_tensors
        if dpred is None and de3 is None and de4 is None:
            return (None,) * 12
        B, _, _, C2 = e2.shape
        C3, NB = w3.shape[0], w14.shape[0]
        dpred = None if dpred is None else _c(dpred)
        de3 = None if de3 is None else _c(de3)
        de4 = None if de4 is None else _c(de4)
        want_w = any(ctx.needs_input_grad[4:])
        bufs, rets = zip(*[_gbuf(prm, want_w) for prm in ctx.params])
        de2 = torch.empty_like(e2) if ctx.needs_input_grad[0] else None

        def launch(grads, dx):
            g = [_p(t) for t in grads] if grads is not None else [None] * 8
            _call("cgs_tail_bwd", _p(e2), _p(m_e2), _p(m_e3), _p(m_v), _p(w3), _p(w14), _p(w1), _p(w2), _p(e3),
                  _p(idx3, torch.uint8), _p(e4), _p(v), _p(pred), _p(dpred), _p(de3), _p(de4), B, C2, C3, NB,
                  *g, _p(dx), _stream())
        direct = want_w and all(r is None for r in rets)
        if direct and async_wgrad and de2 is not None:
            # critical path: only the input gradient; all weight gradients on the side stream (the head chain is
            # recomputed there, it is < 1 % of the step)
            _fork_wgrad(lambda: launch(bufs, None), (e2, m_e2, m_e3, m_v, e3, idx3, e4, v, pred, dpred, de3, de4), True)
            launch(None, de2)
        else:
            launch(bufs if want_w else None, de2)
        return (de2, None, None, None) + tuple(rets)


class DecBlock(torch.autograd.Function):
    """cat(skip, nearest_up(up, 2**shift)) -> Conv2d(3,1,1) [-> LeakyReLU(0.01)]: one UnetDecoder
    stage (reference nets.py:503-521).  skip [B,H,W,C0], up [B,H>>shift,W>>shift,C1]."""

    @staticmethod
    def forward(ctx, skip, up, w, b, shift, leaky):
        B, H, W, C0 = skip.shape
        C1 = up.shape[3]
        Cout = w.shape[0]
        out = torch.empty((B, H, W, Cout), device=skip.device, dtype=torch.float32)
        conv3x3(_src(SRC_CATUP, C0 + C1, skip, up, C0=C0, shift=shift), w, b, B, H, W, Cout,
                EPI_LEAKY if leaky else EPI_LINEAR, out)
        ctx.save_for_backward(skip, up, w, out if leaky else None)
        ctx.shift, ctx.leaky, ctx.params = shift, leaky, (w, b)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        skip, up, w, out = ctx.saved_tensors
        if dout is None:
            return (None,) * 6
        dout = _c(dout)
        B, H, W, C0 = skip.shape
        C1 = up.shape[3]
        Cout = w.shape[0]
        shift = ctx.shift
        dy = _src(SRC_LEAKYGRAD, Cout, dout, out) if ctx.leaky else _src(SRC_PLAIN, Cout, dout)
        dskip = dup = rw = rb = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dw, rw = _gbuf(ctx.params[0], True)
            db, rb = _gbuf(ctx.params[1], True)
            xs = _src(SRC_CATUP, C0 + C1, skip, up, C0=C0, shift=shift)
            _fork_wgrad(lambda: wgrad3x3(xs, dy, B, H, W, dw, db), (skip, up, dout, out), rw is None and rb is None)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            if ctx.needs_input_grad[0]:
                dskip = torch.empty_like(skip)
            if ctx.needs_input_grad[1]:
                dup = torch.zeros_like(up) if shift == 2 else torch.empty_like(up)
            conv3x3(dy, w, None, B, H, W, C0 + C1, EPI_SPLIT_UP, dskip, transposed=True, out2=dup,
                    C0=C0, shift2=shift)
        return dskip, dup, rw, rb, None, None


class MaskHead(torch.autograd.Function):
    """Conv2d(16,1,3,1,1) -> Sigmoid (reference nets.py:490-491) [+ fused >= threshold, main.py:1164]."""

    @staticmethod
    def forward(ctx, m, w, b, thresh):
        B, H, W, Cm = m.shape
        z = torch.empty((B, H, W, 1), device=m.device, dtype=torch.float32)
        hard = None
        if thresh is not None:
            hard = torch.empty((B, H, W, 1), device=m.device, dtype=torch.uint8)
        conv3x3(_src(SRC_PLAIN, Cm, m), w, b, B, H, W, 1, EPI_SIGMOID, z, idx_out=hard,
                thresh=float(thresh) if thresh is not None else 0.0)
        ctx.save_for_backward(m, w, z)
        ctx.params = (w, b)
        ctx.set_materialize_grads(False)
        if hard is None:
            hard = torch.empty(0, device=m.device, dtype=torch.uint8)
        ctx.mark_non_differentiable(hard)
        return z, hard

    @staticmethod
    def backward(ctx, dz, _dhard):
        m, w, z = ctx.saved_tensors
        if dz is None:
            return None, None, None, None
        dz = _c(dz)
        B, H, W, Cm = m.shape
        dy = _src(SRC_SIGGRAD, 1, dz, z)
        dm = rw = rb = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, rw = _gbuf(ctx.params[0], True)
            db, rb = _gbuf(ctx.params[1], True)
            xs = _src(SRC_PLAIN, Cm, m)
            _fork_wgrad(lambda: wgrad3x3(xs, dy, B, H, W, dw, db), (m, dz, z), rw is None and rb is None)
        if ctx.needs_input_grad[0]:
            dm = torch.empty_like(m)
            conv3x3(dy, w, None, B, H, W, Cm, EPI_LINEAR, dm, transposed=True)
        return dm, rw, rb, None


class Dense(torch.autograd.Function):
    """UnetDecoder.dec[4]: Conv2d(32c,32c,1) on the 1x1 bottleneck (reference nets.py:484,500-501)."""

    @staticmethod
    def forward(ctx, x, w, b):
        B = x.shape[0]
        N, K = w.shape[0], w.shape[1]
        out = torch.empty((B, 1, 1, N), device=x.device, dtype=torch.float32)
        _call("cgs_dense_fwd", _p(x), _p(w), _p(b), B, K, N, _p(out), _stream())
        ctx.save_for_backward(x, w)
        ctx.params = (w, b)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        if dout is None:
            return None, None, None
        dout = _c(dout)
        B = x.shape[0]
        N, K = w.shape[0], w.shape[1]
        want_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, rw = _gbuf(ctx.params[0], want_w)
        db, rb = _gbuf(ctx.params[1], want_w)
        _call("cgs_dense_bwd", _p(x), _p(w), _p(dout), B, K, N, _p(dx), _p(dw), _p(db), _stream())
        return dx, rw, rb


class Occlude(torch.autograd.Function):
    """a*(1-z) + z*b with z broadcast over channels (reference main.py:395,406).  NHWC tensors."""

    @staticmethod
    def forward(ctx, a, b, z):
        out = torch.empty_like(a)
        npix = a.numel() // a.shape[-1]
        _call("cgs_occlude_fwd", _p(a), _p(b), _p(z), npix, a.shape[-1], _p(out), _stream())
        ctx.save_for_backward(a, b, z)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, z = ctx.saved_tensors
        if g is None:
            return None, None, None
        g = _c(g)
        npix = a.numel() // a.shape[-1]
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        dz = torch.empty_like(z) if ctx.needs_input_grad[2] else None
        _call("cgs_occlude_bwd", _p(a), _p(b), _p(z), _p(g), npix, a.shape[-1], _p(dz), _p(da), _p(db), _stream())
        return da, db, dz


class PredLoss(torch.autograd.Function):
    """F.mse_loss(pred, target) / F.binary_cross_entropy (reference main.py:193-195,400,411), mean reduction."""

    @staticmethod
    def forward(ctx, pred, target, bce):
        p = _c(pred.reshape(-1))
        t = _c(target.reshape(-1).to(torch.float32))
        loss = torch.empty(1, device=p.device, dtype=torch.float32)
        grad = torch.empty_like(p)
        _call("cgs_pred_loss", _p(p), _p(t), p.numel(), int(bce), 1.0, _p(loss), _p(grad), _stream())
        ctx.save_for_backward(grad)
        ctx.shape = pred.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).reshape(ctx.shape), None, None


class MaskReg(torch.autograd.Function):
    """L1*l1_loss(vf*Z,0) + L2*mse_loss(vf*Z,0) (reference main.py:415-429); vf = 1 or 1-pred[frame]."""

    @staticmethod
    def forward(ctx, z, vpred, l1, l2):
        zc = _c(z)
        n = zc.numel()
        per_frame = n // zc.shape[0]
        loss = torch.empty(1, device=z.device, dtype=torch.float32)
        grad = torch.empty_like(zc)
        vp = None if vpred is None else _c(vpred.detach().reshape(-1))
        _call("cgs_mask_reg", _p(zc), _p(vp), n, per_frame, float(l1), float(l2), 1.0, _p(loss), _p(grad), _stream())
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


def frames_to_float(frames_u8, roll=0, roll_dev=None):
    """uint8 NHWC [B,H,W,C] (device) -> fp32 NHWC /255 with circular W-roll (main.py:189,584-591).
    roll > 0 == `cat(X[:, :, roll:], X[:, :, :roll])`; roll < 0 == the `-xshift` branch."""
    if not frames_u8.is_cuda:
        raise CgsError("frames_to_float needs a CUDA tensor")
    x = _c(frames_u8)
    B, H, W, Cc = x.shape
    out = torch.empty((B, H, W, Cc), device=x.device, dtype=torch.float32)
    _call("cgs_frames_to_float", _p(x, torch.uint8), B, H, W, Cc, int(roll), _p(roll_dev, torch.int32), _p(out), _stream())
    return out


def dropout_masks(shapes, p, seed, state):
    """Multiplicative dropout masks (0 or 1/(1-p)) for all `shapes` from ONE kernel launch; returns views of one buffer.
    `state`: int64 device tensor [2] (call counter, ticket) owned by the module."""
    sizes = [int(torch.Size(s).numel()) for s in shapes]
    offs, tot = [], 0
    for n in sizes:
        offs.append(tot)
        tot += (n + 3) & ~3          # keep every view 16-byte aligned
    buf = torch.empty(tot, device=state.device, dtype=torch.float32)
    _call("cgs_dropout_masks", _p(buf), tot, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(state, torch.int64), _stream())
    return [buf[o:o + n].view(s) for o, n, s in zip(offs, sizes, shapes)]


def critic_fused_supported(critic):
    """True when the whole-step kernel covers this NewCritic (chfak=1 geometry) and the tensor-core mode is on."""
    f = critic.features
    return bool(_precision and _lib.lib().cgs_critic_fused_supported(
        f[0].out_channels, f[3].out_channels, f[6].out_channels, f[10].out_channels, f[14].out_channels))


def _critic_bucket_span(params):
    """(FlatAdam, offset) if the 14 critic parameters are one contiguous run, in registration order, of a FlatAdam's
    flat gradient bucket (then the kernel may hand its gradient over as per-CTA partial vectors); else (None, 0)."""
    opt = getattr(params[0], "_cgs_opt", None)
    if opt is None:
        return None, 0
    base = opt.gflat.data_ptr()
    off0 = None
    expect = None
    for q in params:
        if getattr(q, "_cgs_opt", None) is not opt or q.grad is not getattr(q, "_cgs_grad", None):
            return None, 0
        o = (q.grad.data_ptr() - base) // 4
        if off0 is None:
            off0 = expect = o
        if o != expect:
            return None, 0
        expect += q.numel()
    return opt, off0


def critic_train_fused(critic, frames_u8, target, roll=0, masks=(None, None, None), loss_grad=1.0, bce=False,
                       use_partials=True, rng=None, fuse_adam=False, bf16=False):
    """Forward + loss + backward of one critic_pipe step (reference main.py:185-198) in ONE kernel
    (cgs_critic_train_fused).  The parameter gradient is ADDED to the parameters' `.grad`: by REDs, or — when the
    parameters sit contiguously in a FlatAdam bucket — as per-CTA partial vectors that `FlatAdam.step()` sums inside the
    Adam kernel (no atomics, bit-reproducible).  Returns (loss scalar tensor, pred [B])."""
    B = frames_u8.shape[0]
    params = list(critic.parameters())          # registration order == state_dict order (nets.py:169-195)
    assert len(params) == 14
    L = _lib.lib()
    opt, off = _critic_bucket_span(params) if use_partials else (None, 0)
    w = _lib.CriticWeights(*[_p(q.detach()) for q in params])
    pred = torch.empty(B, device=frames_u8.device, dtype=torch.float32)
    loss = torch.empty(1, device=frames_u8.device, dtype=torch.float32)
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    m2, m3, mv = masks
    # rng = (p, seed, int64 state tensor): masks drawn inside the kernel from the module's Philox stream
    rp, rseed, rstate = (float(rng[0]), int(rng[1]) & 0xFFFFFFFFFFFFFFFF, _p(rng[2], torch.int64)) if rng is not None else (0.0, 0, None)
    if opt is not None:
        grid, stride = L.cgs_critic_fused_grid(B), L.cgs_critic_fused_partial_stride()
        opt.flush_partials()                                       # an unconsumed earlier hand-over goes into the bucket first
        buf = opt.partial_buffer(grid * stride)
        nparam = sum(q.numel() for q in params)
        adam = opt.fused_adam_args(off, nparam) if fuse_adam else None       # single GPU, bucket == the critic
        if bf16:            # bf16 tensor-core operands (csrc/hg_critic.cu); same partial-vector / Adam / all-reduce contract
            _call("cgs_critic_train_bf16", _p(frames_u8, torch.uint8), _p(target), B, r, rd, _p(m2), _p(m3), _p(mv),
                  rp, rseed, rstate, C.byref(w), _p(buf), C.byref(adam) if adam is not None else None, float(loss_grad),
                  int(bool(bce)), _p(pred), _p(loss), _stream())
        else:
            _call("cgs_critic_train_fused", _p(frames_u8, torch.uint8), _p(target), B, r, rd, _p(m2), _p(m3), _p(mv),
                  rp, rseed, rstate, C.byref(w), None, _p(buf), C.byref(adam) if adam is not None else None, float(loss_grad),
                  int(bool(bce)), _p(pred), _p(loss), _stream())
        if adam is not None:
            global _weights_epoch
            _weights_epoch += 1
            opt.adam_done_in_kernel = True          # the coming opt.step() has nothing left to do
        else:
            opt.pending_partials = (buf, grid, stride, off, nparam)
        return loss.reshape(()), pred
    if bf16:
        raise CgsError("critic_train_fused(bf16=True) hands its gradient over as partial vectors: the critic's parameters must sit "
                       "contiguously in a FlatAdam bucket")
    grads = []
    for q in params:
        if q.grad is None:
            q.grad = torch.zeros_like(q)
        o = getattr(q, "_cgs_opt", None)
        if o is not None and q.grad is getattr(q, "_cgs_grad", None):
            o._clean = False
        grads.append(q.grad)
    g = _lib.CriticWeights(*[_p(t) for t in grads])
    _call("cgs_critic_train_fused", _p(frames_u8, torch.uint8), _p(target), B, r, rd, _p(m2), _p(m3), _p(mv),
          rp, rseed, rstate, C.byref(w), C.byref(g), None, None, float(loss_grad), int(bool(bce)), _p(pred), _p(loss), _stream())
    return loss.reshape(()), pred


class CriticLossXGrad(torch.autograd.Function):
    """loss = F.mse_loss(critic(x), target) (or BCE) for a FROZEN critic, with d loss / d x computed in the same kernel
    (cgs_critic_loss_xgrad): forward returns the loss scalar, backward is one multiply."""

    @staticmethod
    def forward(ctx, x, target, critic, masks, rng, bce):
        B = x.shape[0]
        w = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
        pred = torch.empty(B, device=x.device, dtype=torch.float32)
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        m2, m3, mv = masks
        rp, rseed, rstate = (float(rng[0]), int(rng[1]) & 0xFFFFFFFFFFFFFFFF, _p(rng[2], torch.int64)) if rng is not None else (0.0, 0, None)
        _call("cgs_critic_loss_xgrad", _p(x), _p(target), B, _p(m2), _p(m3), _p(mv), rp, rseed, rstate, C.byref(w), 1.0,
              int(bce), _p(pred), _p(loss), _p(dx), _stream())
        ctx.save_for_backward(dx)
        ctx.pred = pred
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None, None, None, None, None


def critic_loss_xgrad(critic, x_nhwc, target, masks=(None, None, None), rng=None, bce=False):
    """Loss of a frozen critic on fp32 NHWC frames [B,64,64,3] with the input gradient from the same kernel.
    `x_nhwc` may require grad (e.g. the occlusion blend); the critic's parameters get no gradient."""
    return CriticLossXGrad.apply(_c(x_nhwc), _c(target.to(torch.float32)), critic, masks, rng, bce)


def _rng_args(rng):
    return (float(rng[0]), int(rng[1]) & 0xFFFFFFFFFFFFFFFF, _p(rng[2], torch.int64)) if rng is not None else (0.0, 0, None)


def hg_score(critic, A_u8, B_u8, z, target_replace, target_inject=None, roll=0, masks=None, masks_inject=None, rng=None,
             loss_grad=1.0, vpred=None, l1=0.0, l2=0.0):
    """The scored blends of one Hourglass step (reference main.py:395-429) in ONE kernel (cgs_hg_score): uint8 frames A
    (rolled) and B [B,64,64,3], mask z [B,64,64] (any shape with B*4096 elements) -> `replaced`/`injected` blends in shared
    memory -> frozen critic -> MSE against negpred / pred -> backward into the blends, contracted with (B - A) / (A - B),
    plus the L1/L2 mask regulariser.  Returns (losses [4] = replace, inject, L1, L2 terms; dz [B,64,64] = loss_grad *
    d(sum)/dZ; pred_replace [B]; pred_inject [B] or None).  No autograd: the caller feeds dz to the masker's backward."""
    Bn = A_u8.shape[0]
    dev = A_u8.device
    w = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    zc = _c(z.detach())
    assert zc.numel() == Bn * 4096
    pr = torch.empty(Bn, device=dev, dtype=torch.float32)
    pi = torch.empty(Bn, device=dev, dtype=torch.float32) if target_inject is not None else None
    losses = torch.empty(4, device=dev, dtype=torch.float32)
    dz = torch.empty((Bn, 64, 64), device=dev, dtype=torch.float32)
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    m = masks if masks is not None else (None, None, None)
    mi = masks_inject if masks_inject is not None else (None, None, None)
    rp, rseed, rstate = _rng_args(rng)
    _call("cgs_hg_score", _p(A_u8, torch.uint8), _p(B_u8, torch.uint8), Bn, r, rd, _p(zc), _p(_c(target_replace.detach())),
          _p(_c(target_inject.detach())) if target_inject is not None else None, _p(m[0]), _p(m[1]), _p(m[2]),
          _p(mi[0]), _p(mi[1]), _p(mi[2]), rp, rseed, rstate, C.byref(w), float(loss_grad),
          _p(_c(vpred.detach())) if vpred is not None else None, float(l1), float(l2), _p(pr), _p(pi), _p(losses), _p(dz), _stream())
    return losses, dz, pr, pi


def critic_saliency(critic, x_nhwc):
    """The saliency baseline of Handler.eval (reference main.py:945-951): `pred.mean().backward(); batch.grad.abs().sum(1)` for
    fp32 NHWC frames, forward and input gradient in ONE kernel.  Returns (pred [B,1], saliency [B,1,64,64])."""
    x = _c(x_nhwc.detach())
    B = x.shape[0]
    w = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    pred = torch.empty(B, device=x.device, dtype=torch.float32)
    loss = torch.empty(1, device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    dummy = torch.zeros(B, device=x.device, dtype=torch.float32)
    _call("cgs_critic_loss_xgrad", _p(x), _p(dummy), B, None, None, None, 0.0, 0, None, C.byref(w), 1.0, 2, _p(pred), _p(loss),
          _p(dx), _stream())
    return pred.unsqueeze(1), dx.abs().sum(dim=3).unsqueeze(1)


def critic_forward_fused(critic, x_nhwc, masks=(None, None, None), rng=None):
    """pred [B,1] = critic(x) for fp32 NHWC frames in ONE kernel, no autograd (the forward-only variant of
    cgs_critic_loss_xgrad): `negpred = critic(B)` under no_grad, extract_contrastive_data."""
    x = _c(x_nhwc.detach())
    B = x.shape[0]
    w = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    pred = torch.empty(B, device=x.device, dtype=torch.float32)
    m2, m3, mv = masks
    rp, rseed, rstate = (float(rng[0]), int(rng[1]) & 0xFFFFFFFFFFFFFFFF, _p(rng[2], torch.int64)) if rng is not None else (0.0, 0, None)
    _call("cgs_critic_loss_xgrad", _p(x), None, B, _p(m2), _p(m3), _p(mv), rp, rseed, rstate, C.byref(w), 1.0, 0, _p(pred), None,
          None, _stream())
    return pred.unsqueeze(1)


def critic_forward_frames(critic, frames_u8, roll=0, masks=(None, None, None), rng=None):
    """pred [B,1] = critic(frames / 255) on raw uint8 NHWC frames in ONE kernel, no autograd (cgs_critic_forward_frames)."""
    B = frames_u8.shape[0]
    w = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    pred = torch.empty(B, device=frames_u8.device, dtype=torch.float32)
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    m2, m3, mv = masks
    rp, rseed, rstate = _rng_args(rng)
    _call("cgs_critic_forward_frames", _p(frames_u8, torch.uint8), B, r, rd, _p(m2), _p(m3), _p(mv), rp, rseed, rstate, C.byref(w),
          _p(pred), _stream())
    return pred.unsqueeze(1)


def infer_fused_supported(critic, masker):
    """True when the fused encoder+decoder inference kernel covers these modules (chfak=1 geometry, tf32 mode, eval)."""
    f, d = critic.features, masker.dec
    return bool(_precision and not critic.training
                and _lib.lib().cgs_critic_fused_supported(f[0].out_channels, f[3].out_channels, f[6].out_channels,
                                                          f[10].out_channels, f[14].out_channels)
                and tuple(d[3].weight.shape) == (16, 48, 3, 3) and tuple(d[0].weight.shape) == (8, 16, 3, 3)
                and tuple(d[4].weight.shape) == (32, 32, 1, 1) and tuple(masker.masker[0].weight.shape) == (16, 11, 3, 3))


def infer_encode_decode(critic, masker, frames_u8):
    """critic(X, collect=True) + dec[4..0] of the masker on raw uint8 frames in ONE kernel (cgs_infer_fused):
    returns (pred [B,1], o0 [B,32,32,8] NHWC).  The decoder weights are packed into mma fragment order once per
    weight version (cached on the masker)."""
    L = _lib.lib()
    B = frames_u8.shape[0]
    d = masker.dec
    ver = (_weights_epoch,) + tuple((w.data_ptr(), w._version) for w in (d[3].weight, d[2].weight, d[1].weight, d[0].weight))
    cache = getattr(masker, "_cgs_pack", None)
    if cache is None or cache[0] != ver:
        pack = torch.empty(L.cgs_infer_pack_floats(), device=frames_u8.device, dtype=torch.float32)
        _call("cgs_infer_pack_decoder", _p(d[3].weight.detach()), _p(d[2].weight.detach()), _p(d[1].weight.detach()),
              _p(d[0].weight.detach()), _p(pack), _stream())
        masker._cgs_pack = cache = (ver, pack)
    w = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    pred = torch.empty(B, device=frames_u8.device, dtype=torch.float32)
    o0 = torch.empty((B, 32, 32, 8), device=frames_u8.device, dtype=torch.float32)
    _call("cgs_infer_fused", _p(frames_u8, torch.uint8), B, C.byref(w), _p(d[4].weight.detach()), _p(d[4].bias.detach()),
          _p(d[3].bias.detach()), _p(d[2].bias.detach()), _p(d[1].bias.detach()), _p(d[0].bias.detach()), _p(cache[1]),
          _p(pred), _p(o0), _stream())
    return pred.unsqueeze(1), o0


def masker_fused(masker, frames_u8, o0, thresh=None):
    """masker[0..3] on cat(X, ups(o0)) in ONE kernel (cgs_masker_fused): returns (mask [B,1,64,64], hard uint8 or None)."""
    B = frames_u8.shape[0]
    m0, m2 = masker.masker[0], masker.masker[2]
    mask = torch.empty((B, 1, 64, 64), device=frames_u8.device, dtype=torch.float32)
    hard = torch.empty((B, 1, 64, 64), device=frames_u8.device, dtype=torch.uint8) if thresh is not None else None
    _call("cgs_masker_fused", _p(frames_u8, torch.uint8), _p(o0), B, _p(m0.weight.detach()), _p(m0.bias.detach()),
          _p(m2.weight.detach()), _p(m2.bias.detach()), float(thresh if thresh is not None else 0.0), _p(mask),
          _p(hard, torch.uint8), _stream())
    return mask, hard


# ---------------------------------------------------------------------------------------------------------------------
# bf16 whole-frame Hourglass kernels (csrc/hg_forward.cu, csrc/hg_backward.cu)
def hg_supported(critic, masker):
    """True when the whole-frame Hourglass kernels cover these modules: chfak = 1 geometry, tensor-core precision mode."""
    f, d = critic.features, masker.dec
    return bool(_precision
                and _lib.lib().cgs_critic_fused_supported(f[0].out_channels, f[3].out_channels, f[6].out_channels,
                                                          f[10].out_channels, f[14].out_channels)
                and tuple(d[3].weight.shape) == (16, 48, 3, 3) and tuple(d[0].weight.shape) == (8, 16, 3, 3)
                and tuple(d[4].weight.shape) == (32, 32, 1, 1) and tuple(masker.masker[0].weight.shape) == (16, 11, 3, 3))


def _masker_weights(masker):
    return _lib.MaskerWeights(*[_p(q.detach()) for q in masker.parameters()])


def hg_pack(critic, masker, out=None, cached=True):
    """Weight fragments of both networks (cgs_hg_pack).  cached: re-packed only when a weight tensor's version (or the
    global weights epoch, bumped by the optimizer kernels) changes; out: pack into this buffer unconditionally (training:
    one launch per step, CUDA-graph capturable)."""
    L = _lib.lib()
    dev = next(masker.parameters()).device
    if out is None and cached:
        ver = (_weights_epoch,) + tuple((q.data_ptr(), q._version) for q in list(critic.parameters()) + list(masker.parameters()))
        cache = getattr(masker, "_cgs_hg_pack", None)
        if cache is not None and cache[0] == ver:
            return cache[1]
    pack = out if out is not None else torch.empty(L.cgs_hg_pack_words(), device=dev, dtype=torch.int32)
    cw = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    mw = _masker_weights(masker)
    _call("cgs_hg_pack", C.byref(cw), C.byref(mw), _p(pack, torch.int32), _stream())
    if out is None and cached:
        masker._cgs_hg_pack = (ver, pack)
    return pack


def hg_forward(critic, masker, frames_u8, roll=0, train=False, masks=None, rng=None, thresh=None, tape=None, pack=None):
    """critic(X, collect=True) + masker(X, embeds) on raw uint8 frames in ONE kernel (cgs_hg_forward).
    Returns (pred [B,1], mask [B,1,64,64], hard uint8 [B,1,64,64] or None).  train: dropout (masks = forced NHWC triple,
    or rng = critic._dropout_rng()); tape: uint8 [B, cgs_hg_tape_bytes()] buffer for hg_backward."""
    B = frames_u8.shape[0]
    dev = frames_u8.device
    if pack is None:
        pack = hg_pack(critic, masker)
    cw = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    mw = _masker_weights(masker)
    pred = torch.empty(B, device=dev, dtype=torch.float32)
    mask = torch.empty((B, 1, 64, 64), device=dev, dtype=torch.float32)
    hard = torch.empty((B, 1, 64, 64), device=dev, dtype=torch.uint8) if thresh is not None else None
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    m = masks if masks is not None else (None, None, None)
    rp, rseed, rstate = _rng_args(rng)
    _call("cgs_hg_forward", _p(frames_u8, torch.uint8), B, r, rd, C.byref(cw), C.byref(mw), _p(pack, torch.int32), int(bool(train)),
          _p(m[0]), _p(m[1]), _p(m[2]), rp, rseed, rstate, float(thresh if thresh is not None else 0.0), _p(pred), _p(mask),
          _p(hard, torch.uint8), _p(tape, torch.uint8), _stream())
    return pred.unsqueeze(1), mask, hard


def hg_score_bf16(critic, A_u8, B_u8, z, pack, target_replace=None, target_inject=None, roll=0, masks=None, rng=None,
                  loss_grad=1.0, vpred=None, l1=0.0, l2=0.0):
    """The critic-scoring part of one frozen-critic Hourglass step in ONE bf16 kernel (cgs_hg_score_bf16; reference
    main.py:365-367, 395-429).  target_replace None: the kernel first computes negpred = critic(B) itself.  masks: None, or the
    forced dropout triples (critic(B), critic(replaced), critic(injected)) - entries of passes that do not run may be None.
    Returns (losses [4] = replace, inject, L1, L2; dz [B,64,64]; negpred [B] or None; pred_replace [B]; pred_inject [B] or None)."""
    Bn = A_u8.shape[0]
    dev = A_u8.device
    w = _lib.CriticWeights(*[_p(q.detach()) for q in critic.parameters()])
    zc = _c(z.detach())
    assert zc.numel() == Bn * 4096
    new = lambda: torch.empty(Bn, device=dev, dtype=torch.float32)
    neg = new() if target_replace is None else None
    pr = new()
    pi = new() if target_inject is not None else None
    losses = torch.empty(4, device=dev, dtype=torch.float32)
    dz = torch.empty((Bn, 64, 64), device=dev, dtype=torch.float32)
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    m9 = None
    if masks is not None and any(m is not None and m[0] is not None for m in masks):
        flat = []
        for m in masks:
            flat += [None, None, None] if (m is None or m[0] is None) else [_p(t) for t in m]
        m9 = (C.c_void_p * 9)(*flat)
    rp, rseed, rstate = _rng_args(rng)
    _call("cgs_hg_score_bf16", _p(A_u8, torch.uint8), _p(B_u8, torch.uint8), Bn, r, rd, _p(zc),
          _p(_c(target_replace.detach())) if target_replace is not None else None,
          _p(_c(target_inject.detach())) if target_inject is not None else None, m9, rp, rseed, rstate, C.byref(w),
          _p(pack, torch.int32), float(loss_grad), _p(_c(vpred.detach())) if vpred is not None else None, float(l1), float(l2),
          _p(neg), _p(pr), _p(pi), _p(losses), _p(dz), _stream())
    return losses, dz, neg, pr, pi


def hg_tape(B, device):
    return torch.empty((B, _lib.lib().cgs_hg_tape_bytes()), device=device, dtype=torch.uint8)


def hg_backward(masker, frames_u8, tape, mask, dz, roll=0, pack=None, partials=None, debug=None):
    """The masker's whole backward in ONE kernel (cgs_hg_backward): returns (partials [grid, stride], grid)."""
    L = _lib.lib()
    B = frames_u8.shape[0]
    grid, stride = L.cgs_hg_grid(B), L.cgs_hg_partial_stride()
    if partials is None:
        partials = torch.empty((grid, stride), device=frames_u8.device, dtype=torch.float32)
    assert partials.numel() >= grid * stride
    mw = _masker_weights(masker)
    rd, r = (_p(roll, torch.int32), 0) if torch.is_tensor(roll) else (None, int(roll or 0))
    _call("cgs_hg_backward", _p(frames_u8, torch.uint8), B, r, rd, C.byref(mw), _p(pack, torch.int32), _p(tape, torch.uint8),
          _p(_c(mask.detach())), _p(_c(dz.detach())), _p(partials), _p(debug), _stream())
    return partials, grid


def reduce_partials(g, buf, n_partials, stride, offset, length):
    _call("cgs_reduce_partials", _p(g), g.numel(), _p(buf), int(n_partials), int(stride), int(offset), int(length), _stream())


def adam_step_partials(p, g, m, v, step_state, buf, n_partials, stride, offset, length, lr=1e-3, betas=(0.9, 0.999),
                       eps=1e-8, grad_scale=1.0):
    """adam_step whose gradient is g + the sum of the per-CTA partial vectors in `buf`; g is cleared."""
    global _weights_epoch
    _weights_epoch += 1
    _call("cgs_adam_step_partials", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(betas[0]), float(betas[1]),
          float(eps), _p(step_state, torch.int32), float(grad_scale), _p(buf), int(n_partials), int(stride), int(offset),
          int(length), _stream())


def threshold(z, thresh, strict=False):
    zc = _c(z)
    hard = torch.empty(zc.shape, device=z.device, dtype=torch.uint8)
    _call("cgs_threshold", _p(zc), zc.numel(), float(thresh), int(strict), _p(hard, torch.uint8), _stream())
    return hard


def iou_counts(z, gt_u8, thresh, counts, strict=True):
    """counts (int64 [2] device tensor, caller-zeroed) += (#(hard & gt), #(hard | gt)), hard = z > thresh (reference
    main.py:964) or >= when not strict; the reference's get_iou (main.py:1265-1270) is counts[0] / counts[1]."""
    zc, gc = _c(z), _c(gt_u8)
    assert zc.numel() == gc.numel()
    _call("cgs_iou_counts", _p(zc), _p(gc, torch.uint8), zc.numel(), float(thresh), int(strict), _p(counts, torch.int64), _stream())
    return counts


def adam_step(p, g, m, v, step_state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, clear_grad=False):
    """Flat-bucket Adam (torch.optim.Adam defaults, main.py:178).  step_state: int32 device tensor [2] = (steps applied,
    ticket); the kernel advances it.  clear_grad: zero g in the same pass (fused zero_grad)."""
    global _weights_epoch
    _weights_epoch += 1
    _call("cgs_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(betas[0]), float(betas[1]),
          float(eps), _p(step_state, torch.int32), float(grad_scale), int(clear_grad), _stream())


def p2p_stage(g, npad, sym, step_state, partials=None):
    """Local gradient (bucket + optional per-CTA partials) -> this rank's symmetric buffer slot of the coming step."""
    buf, rows, stride, off, length = partials if partials is not None else (None, 0, 0, 0, 0)
    _call("cgs_p2p_stage", _p(g), g.numel(), int(npad), _p(sym), _p(buf), int(rows), int(stride), int(off), int(length),
          _p(step_state, torch.int32), _stream())


def p2p_allreduce_adam(p, m, v, npad, peer_bufs, peer_flags, rank, world, step_state, err, lr=1e-3, betas=(0.9, 0.999),
                       eps=1e-8, grad_scale=1.0):
    """One-shot all-reduce over NVLink peer memory fused with Adam (cgs_p2p_allreduce_adam)."""
    global _weights_epoch
    _weights_epoch += 1
    _call("cgs_p2p_allreduce_adam", _p(p), _p(m), _p(v), p.numel(), int(npad), peer_bufs, peer_flags, int(rank), int(world),
          float(lr), float(betas[0]), float(betas[1]), float(eps), _p(step_state, torch.int32), float(grad_scale),
          _p(err, torch.int32), _stream())


def occlude(a, b, z):
    return Occlude.apply(a, b, z)


def pred_loss(pred, target, bce=False):
    return PredLoss.apply(pred, target, bce)


def mask_reg(z, vpred=None, l1=0.0, l2=0.0):
    return MaskReg.apply(z, vpred, l1, l2)


# ---------------------------------------------------------------------------------------------------------------------
# formats either side of the path (csrc/edges.cu)
def gather_frames(dataset_u8, idx_i32, out=None):
    """out[i] = dataset[idx[i]]: the contrastive batches of main.py:345-353 from a device-resident uint8 dataset."""
    n = idx_i32.numel()
    if out is None:
        out = torch.empty((n,) + tuple(dataset_u8.shape[1:]), device=dataset_u8.device, dtype=torch.uint8)
    assert dataset_u8[0].numel() == 12288 and out.numel() == n * 12288
    _call("cgs_gather_frames", _p(dataset_u8, torch.uint8), dataset_u8.shape[0], _p(idx_i32, torch.int32), n, _p(out, torch.uint8), _stream())
    return out


_FRAME_LUT = {}


def _frame_lut(device):
    """(b / 255.0 * 255).astype(uint8) for b = 0..255, computed exactly as the reference does (float64, main.py:1127, 1216)."""
    if device not in _FRAME_LUT:
        import numpy as np
        _FRAME_LUT[device] = torch.from_numpy(((np.arange(256, dtype=np.uint8) / 255.0) * 255).astype(np.uint8)).to(device)
    return _FRAME_LUT[device]


def mask_images(mask, hard, frames_u8=None, concatenated=False):
    """PNG-ready uint8 images of main.py:1212-1223: (raw [B,64,64,3], thresholded [B,64,64,3]), or the `-concatenated` strip
    [B,64,192,3] = frame | raw-mask | thresholded-mask."""
    B = mask.shape[0]
    dev = mask.device
    m, h = _c(mask.detach()).reshape(B, 64, 64), _c(hard).reshape(B, 64, 64)
    if concatenated:
        strip = torch.empty((B, 64, 192, 3), device=dev, dtype=torch.uint8)
        _call("cgs_mask_images", _p(m), _p(h, torch.uint8), B, _p(frames_u8, torch.uint8), _p(_frame_lut(dev), torch.uint8), 1,
              _p(strip, torch.uint8), None, _stream())
        return strip
    raw = torch.empty((B, 64, 64, 3), device=dev, dtype=torch.uint8)
    thr = torch.empty((B, 64, 64, 3), device=dev, dtype=torch.uint8)
    _call("cgs_mask_images", _p(m), _p(h, torch.uint8), B, None, None, 0, _p(raw, torch.uint8), _p(thr, torch.uint8), _stream())
    return raw, thr


def saliency_normalize(sal, pred, thresh, global_norm=False):
    """The saliency baseline's normalisation and threshold (main.py:974-993): returns (salM [B,1,64,64], salhardM uint8)."""
    B = sal.shape[0]
    s = _c(sal.detach()).reshape(B, 4096)
    pr = _c(pred.detach()).reshape(B)
    out = torch.empty((B, 1, 64, 64), device=s.device, dtype=torch.float32)
    hard = torch.empty((B, 1, 64, 64), device=s.device, dtype=torch.uint8)
    gn = None
    if global_norm:                                     # norm = (salM * (salM >= 0)).mean() * thresh, main.py:978
        gn = ((s * (s >= 0)).mean() * thresh).reshape(1).float()
    _call("cgs_saliency_normalize", _p(s), _p(pr), B, int(64 * 64 * thresh), float(thresh), _p(gn), _p(out), _p(hard, torch.uint8), None,
          _stream())
    return out, hard
