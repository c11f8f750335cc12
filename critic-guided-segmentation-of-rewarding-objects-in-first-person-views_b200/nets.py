"""Drop-in replacements for the two live network classes of the reference `nets.py`:
`NewCritic` (reference nets.py:160-212) and `UnetDecoder` (nets.py:452-523).

Same constructor signatures, `forward()` signatures, return structure, parameter
registration order and `state_dict()` keys/shapes (OIHW fp32), so reference
checkpoints load here and vice versa, `Adam(critic.parameters())` sees the same
tensors in the same order, and default initialisation consumes the torch RNG
identically.  The `nn.Conv2d` / `nn.Linear` objects are parameter containers only:
`forward()` dispatches to the fused sm_100a kernels of libcgs_b200.so via
`cgs_b200.ops`.  CUDA only: calling forward on CPU tensors raises (no fallback).
"""
from collections import deque

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import CgsError


def _nhwc(X, C, width, who):
    if X.dim() != 4 or X.shape[1] != C or X.shape[2] != width or X.shape[3] != width:
        raise ValueError(f"{who}: expected input [B,{C},{width},{width}], got {tuple(X.shape)}")
    if X.shape[0] == 0:
        raise ValueError(f"{who}: empty batch")
    if not X.is_cuda:
        raise CgsError(f"{who}: input is on {X.device}; cgs_b200 has no CPU path")
    if X.dtype != torch.float32:
        raise ValueError(f"{who}: expected float32 input (the reference casts with .float()), got {X.dtype}")
    return X.permute(0, 2, 3, 1).contiguous()   # zero-copy for channels_last inputs (main.py:189)


def _nchw(t):
    return t.permute(0, 3, 1, 2)


class NewCritic(nn.Module):
    def __init__(self, width=64, dims=[8, 8, 8, 16], bottleneck=32, colorchs=3, chfak=1, activation=nn.ReLU,
                 pool="max", dropout=0.5):
        super().__init__()
        if pool != "max" or activation is not nn.ReLU:
            raise NotImplementedError("cgs_b200.NewCritic accelerates the live configuration only: "
                                      "pool='max', activation=nn.ReLU (reference main.py:108)")
        if width != 64 or len(dims) != 4:
            raise NotImplementedError("cgs_b200.NewCritic: width must be 64 with four stages "
                                      "(the 4x4 bottleneck conv hard-wires it, reference nets.py:184)")
        self.width = width
        self.colorchs = colorchs
        self.p = float(dropout)
        ch = [int(d) for d in np.array(dims) * chfak]
        self.pool = nn.MaxPool2d(2)
        layers = []
        cin = colorchs
        for i, c in enumerate(ch):
            layers += [nn.Conv2d(cin, c, 3, 1, 1), activation(), self.pool]
            if i >= 2:
                layers.append(nn.Dropout(dropout))
            cin = c
        layers += [nn.Conv2d(cin, bottleneck * chfak, 4), activation()]
        self.features = nn.Sequential(*layers)
        nb = chfak * bottleneck
        self.crit = nn.Sequential(nn.Flatten(), nn.Linear(nb, nb), activation(), nn.Dropout(dropout),
                                  nn.Linear(nb, 1), nn.Sigmoid())
        self._conv_idx = (0, 3, 6, 10)
        self.fuse_tail = True          # features[9..15] + crit in one kernel each way when the shapes fit shared memory
        self.fuse_frame_cast = True    # raw uint8 frames go straight into the first conv / its wgrad (no fp32 copy of the batch)
        self._rng_state = None
        self._rng_seed = 0
        NewCritic._count = getattr(NewCritic, "_count", 0) + 1
        self._instance = NewCritic._count          # distinct Philox stream per module (Handler pins it per role) ...
        self._rank = 0                             # ... and per data-parallel rank
        self._forced_masks = None   # test hook: (m_e2 [B,8,8,8c], m_e3 [B,4,4,16c], m_v [B,32c]) NHWC

    def _dropout_masks(self, B, device):
        if self._forced_masks is not None:
            if isinstance(self._forced_masks, deque):     # one entry per forward call, in call order
                return self._forced_masks.popleft()
            return self._forced_masks
        if not self.training or self.p <= 0.0:
            return None, None, None
        c2 = self.features[6].out_channels
        c3 = self.features[10].out_channels
        nb = self.crit[1].out_features
        if self.p >= 1.0:
            z = lambda *s: torch.zeros(s, device=device, dtype=torch.float32)
            return z(B, 8, 8, c2), z(B, 4, 4, c3), z(B, nb)
        self._ensure_rng(device)
        return tuple(ops.dropout_masks([(B, 8, 8, c2), (B, 4, 4, c3), (B, nb)], self.p, self._rng_seed, self._rng_state))

    def _ensure_rng(self, device):
        """The module's Philox stream state {call counter, ticket}: created once per device (keyed by torch's seed, so
        torch.manual_seed makes it reproducible) and advanced on the device by every kernel that draws from it.  `device` may be
        a string or an index-less torch.device("cuda"): it is normalised before comparing, or the state would be re-created
        (and the stream restarted) on every call."""
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self._rng_state is None or self._rng_state.device != dev:
            self._rng_state = torch.zeros(2, dtype=torch.int64, device=dev)
            self._rng_seed = self._philox_key()

    def _philox_key(self):
        return (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._instance + 0x632BE59BD9B4E019 * self._rank) & 0x7FFFFFFFFFFFFFFF

    def _dropout_rng(self, device):
        """(p, seed, state) of this module's Philox stream when the masks can be drawn inside a fused kernel (train mode,
        0 < p < 1, no forced masks): the same stream `_dropout_masks` would consume, one call per forward.  Else None."""
        if self._forced_masks is not None or not self.training or not (0.0 < self.p < 1.0):
            return None
        self._ensure_rng(device)
        return self.p, self._rng_seed, self._rng_state

    def forward(self, X, collect=False):
        return self._run(_nhwc(X, self.colorchs, self.width, "NewCritic"), None, collect)

    def forward_frames(self, X_u8, roll=0, collect=False):
        """forward() on raw uint8 NHWC frames [B,64,64,3]: `X.permute(0,3,1,2).float()/255.0` and the
        shift_batch roll (reference main.py:185-189, 584-591) are fused into the first conv's operand load."""
        if X_u8.dtype != torch.uint8 or X_u8.dim() != 4 or tuple(X_u8.shape[1:]) != (self.width, self.width, self.colorchs):
            raise ValueError(f"NewCritic.forward_frames: expected uint8 [B,{self.width},{self.width},{self.colorchs}]")
        if not X_u8.is_cuda:
            raise CgsError("NewCritic.forward_frames: input is on CPU; cgs_b200 has no CPU path")
        if self.fuse_frame_cast:
            return self._run(X_u8.contiguous(), roll, collect)       # cast + roll inside features.0's operand load
        if torch.is_tensor(roll):
            return self._run(ops.frames_to_float(X_u8, 0, roll), None, collect)
        return self._run(ops.frames_to_float(X_u8, roll), None, collect)

    def _run(self, x, roll, collect):
        m_e2, m_e3, m_v = self._dropout_masks(x.shape[0], x.device)
        f = self.features
        e0 = ops.EncBlock.apply(x, None, f[0].weight, f[0].bias, roll)
        e1 = ops.EncBlock.apply(e0, None, f[3].weight, f[3].bias)
        e2 = ops.EncBlock.apply(e1, None, f[6].weight, f[6].bias)
        if self.fuse_tail and ops.tail_supported(x.shape[0], e2.shape[3], f[10].out_channels, f[14].out_channels):
            pred, e3, e4 = ops.Tail.apply(e2, m_e2, m_e3, m_v, f[10].weight, f[10].bias, f[14].weight, f[14].bias,
                                          self.crit[1].weight, self.crit[1].bias, self.crit[4].weight, self.crit[4].bias)
        else:
            e3 = ops.EncBlock.apply(e2, m_e2, f[10].weight, f[10].bias)
            pred, e4 = ops.Head.apply(e3, m_e3, m_v, f[14].weight, f[14].bias, self.crit[1].weight, self.crit[1].bias,
                                      self.crit[4].weight, self.crit[4].bias)
        if collect:
            return pred, [_nchw(e0), _nchw(e1), _nchw(e2), _nchw(e3), _nchw(e4)]
        return pred


class Critic(nn.Module):
    """The legacy critic of the reference (nets.py:133-157): four Conv2d(3,1,1) + ReLU + MaxPool2d(2) stages and a 4x4
    valid convolution to ONE channel (+ optional `end` modules, e.g. [nn.Sigmoid()]).  Same constructor, `enc` Sequential layout and
    state_dict keys; forward() runs on the same kernels as NewCritic: `ops.EncBlock` x 4 and the 4x4 convolution on the 4x4 map as
    the dense kernel over the NHWC-flattened map (its weight viewed in (y, x, c) order).  Returns [B, 1, 1, 1] like the reference."""

    def __init__(self, width=64, enc_dim=1, colorchs=3, chfak=1, activation=nn.ReLU, end=[], pool="max"):
        super().__init__()
        if pool != "max" or activation is not nn.ReLU or width != 64:
            raise NotImplementedError("cgs_b200.Critic accelerates pool='max', activation=nn.ReLU, width=64 (the reference's defaults)")
        self.width, self.colorchs = width, colorchs
        mp = nn.MaxPool2d(2)
        modules = [nn.Conv2d(colorchs, 8 * chfak, 3, 1, 1), activation(), mp,
                   nn.Conv2d(8 * chfak, 8 * chfak, 3, 1, 1), activation(), mp,
                   nn.Conv2d(8 * chfak, 8 * chfak, 3, 1, 1), activation(), mp,
                   nn.Conv2d(8 * chfak, 16 * chfak, 3, 1, 1), activation(), mp,
                   nn.Conv2d(16 * chfak, 1, 4)]
        modules.extend(end)
        self.enc = nn.Sequential(*modules)
        self._end = list(end)

    def forward(self, X):
        x = _nhwc(X, self.colorchs, self.width, "Critic")
        e = self.enc
        for i in (0, 3, 6, 9):
            x = ops.EncBlock.apply(x, None, e[i].weight, e[i].bias)
        B = x.shape[0]
        w = e[12].weight                                               # [1, C, 4, 4] -> K order (y, x, c) of the NHWC map
        out = ops.Dense.apply(x.reshape(B, 1, 1, -1), w.permute(0, 2, 3, 1).reshape(1, -1).contiguous(), e[12].bias)
        out = out.reshape(B, 1, 1, 1)
        for m in self._end:
            out = m(out)
        return out


class UnetDecoder(nn.Module):
    def __init__(self, width=64, edims=[8, 8, 8, 16], ddims=[8, 8, 8, 16], bottleneck=32, masker_channels=16,
                 colorchs=3, chfak=1, activation=nn.ReLU, pool="max", upsample=True, pure=False):
        super().__init__()
        if pool != "max" or not upsample or pure:
            raise NotImplementedError("cgs_b200.UnetDecoder accelerates the live configuration only: "
                                      "pool='max', upsample=True, pure=False (reference main.py:109)")
        if width != 64:
            raise NotImplementedError("cgs_b200.UnetDecoder: width must be 64")
        e = [int(v) for v in np.array(edims, dtype=int) * chfak]
        d = [int(v) for v in np.array(ddims, dtype=int) * chfak]
        nb = bottleneck * chfak
        self.width = width
        self.colorchs = colorchs
        self.pool = nn.MaxPool2d(2)
        self.acti = nn.LeakyReLU(0.01)
        self.ups = nn.Upsample(scale_factor=(2, 2))
        self.upsample = upsample
        self.pure = pure
        self.masker_channels = masker_channels
        ins = [e[0] + d[1], e[1] + d[2], e[2] + d[3], e[3] + nb]
        self.dec = [nn.Conv2d(ins[k], d[k], 3, 1, 1) for k in range(4)] + [nn.Conv2d(nb, nb, 1, 1, 0)]
        self.dec_model = nn.Sequential(*self.dec)
        self.masker = nn.Sequential(nn.Conv2d(colorchs + d[0], masker_channels, 3, 1, 1), self.acti,
                                    nn.Conv2d(masker_channels, 1, 3, 1, 1), nn.Sigmoid())

    def _run(self, X, embeds, thresh):
        x = _nhwc(X, self.colorchs, self.width, "UnetDecoder")
        if len(embeds) != 5:
            raise ValueError("UnetDecoder: embeds must be the 5 tensors returned by NewCritic.forward(X, collect=True)")
        e = [t.permute(0, 2, 3, 1).contiguous() for t in embeds]
        dec = self.dec
        o = ops.Dense.apply(e[4], dec[4].weight, dec[4].bias)                          # nets.py:500-501
        o = ops.DecBlock.apply(e[3], o, dec[3].weight, dec[3].bias, 2, False)          # ups(ups()) + cat + dec[3]
        o = ops.DecBlock.apply(e[2], o, dec[2].weight, dec[2].bias, 1, False)
        o = ops.DecBlock.apply(e[1], o, dec[1].weight, dec[1].bias, 1, False)
        o = ops.DecBlock.apply(e[0], o, dec[0].weight, dec[0].bias, 1, False)
        m = ops.DecBlock.apply(x, o, self.masker[0].weight, self.masker[0].bias, 1, True)   # cat(X, ups) + masker[0..1]
        z, hard = ops.MaskHead.apply(m, self.masker[2].weight, self.masker[2].bias, thresh)
        return _nchw(z), (_nchw(hard) if thresh is not None else None)

    def mask_from_o0(self, x_nhwc, o0, thresh=None):
        """The masker half on its own: cat(X, ups(o0)) -> masker[0..3] (reference nets.py:519-523), given dec[0]'s output
        o0 [B,32,32,8] NHWC (e.g. from ops.infer_encode_decode).  x_nhwc: fp32 [B,64,64,3]."""
        m = ops.DecBlock.apply(x_nhwc, o0, self.masker[0].weight, self.masker[0].bias, 1, True)
        z, hard = ops.MaskHead.apply(m, self.masker[2].weight, self.masker[2].bias, thresh)
        return _nchw(z), (_nchw(hard) if thresh is not None else None)

    def forward(self, X, embeds):
        return self._run(X, embeds, None)[0]

    def forward_hard(self, X, embeds, threshold):
        """mask and `mask >= threshold` (uint8) from one fused epilogue (reference main.py:1150,1164)."""
        return self._run(X, embeds, threshold)
