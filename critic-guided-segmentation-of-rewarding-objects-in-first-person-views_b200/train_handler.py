"""Host-side mirror of the reference `Handler` hot loops (reference main.py:66-575, 1103-1167;
TrainHandler.py:74-1096 is the same code): critic regression (`critic_pipe`), pos/neg
split (`extract_contrastive_data`), critic-guided Hourglass training
(`segmentation_training`) and `-process` mask inference (`segment`).

Same flag names, defaults, RNG draw order (torch RNG for shift_batch, numpy RNG for the
contrastive sampling), checkpoint paths and state_dict layout as the reference; the
arithmetic runs in libcgs_b200.so.  Visualisation, MineRL collection, CRF and video
output are out of scope.

Data-parallel: one process per GPU; every rank runs the same loop on its batch shard and the
flat gradient bucket is summed over ranks before the update: inside the whole-step kernel over
NVLink peer memory (critic step), by `cgs_p2p_allreduce_adam` (other buckets), or by an NCCL
all-reduce when symmetric memory is unavailable.  Rank 0's initial parameters are broadcast.
"""
import argparse
import math
import os
from itertools import chain

import numpy as np
import torch

from . import _lib, ops, wide
from .nets import NewCritic, UnetDecoder


def build_parser():
    """The hot-path subset of reference main.py:1462-1533, same names and defaults."""
    p = argparse.ArgumentParser()
    for flag in ("-train", "-frozen", "-noinject", "-separate", "-noevalmode", "-process", "-eval", "-test", "-salience",
                 "-process_salience", "-concatenated"):
        p.add_argument(flag, action="store_true")
    for flag in ("-masker", "-critic", "-cload", "-mload", "-staticnorm", "-salglobal"):
        p.add_argument(flag, type=bool, default=True)
    p.add_argument("--salience-thresh", type=float, default=1.5)
    p.add_argument("--eval-thresh", type=float, default=0.05)
    p.add_argument("--dropout", type=float, default=0.3)
    p.add_argument("--threshrew", type=float, default=0)
    p.add_argument("--datamode", type=str, default="trunk")
    p.add_argument("--chfak", type=int, default=1)
    p.add_argument("--shift", type=int, default=12)
    p.add_argument("--lfak", type=int, default=5)
    p.add_argument("--neck", type=int, default=32)
    p.add_argument("--cepochs", type=int, default=15)
    p.add_argument("--mepochs", type=int, default=1)
    p.add_argument("--high-rew-thresh", type=float, default=0.7)
    p.add_argument("--low-rew-thresh", type=float, default=0.3)
    p.add_argument("--L2", type=float, default=0.0)
    p.add_argument("--L1", type=float, default=0.5)
    p.add_argument("--saveevery", type=int, default=5)
    p.add_argument("--rewidx", type=int, default=1)
    p.add_argument("--testsize", type=int, default=5000)
    p.add_argument("--datasize", type=int, default=100000)
    p.add_argument("--model", type=str, default="default-model")
    p.add_argument("--source-imgs", type=str, default="")
    p.add_argument("--mask-output-imgs", type=str, default="results")
    p.add_argument("--binarymaskthreshold", type=float, default=0.5)
    return p


def parse_args(argv=()):
    args = build_parser().parse_args(list(argv))
    args.live = not args.frozen          # main.py:1537
    args.inject = not args.noinject      # main.py:1538
    args.name = args.model               # main.py:1539
    return args


class FlatAdam:
    """torch.optim.Adam(params) with default hyper-parameters (reference main.py:178, 331-334) over ONE
    flat fp32 bucket: parameters and gradients are re-pointed to views of two flat buffers, so
    `zero_grad` is one memset, the update is one kernel (cgs_adam_step) and the data-parallel
    gradient exchange is one all-reduce."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, world_size=1):
        self.params = [p for p in params]
        assert self.params, "FlatAdam: no parameters"
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.gflat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        # (Adam steps applied, ticket, exchange epoch, -), advanced by the kernels.  The exchange epoch tags the peer-memory
        # all-reduce packets / flags and only ever grows; the step count may be rewound (graph_step._capture)
        self.step_count = torch.zeros(4, device=dev, dtype=torch.int32)
        self._clean = True                                                    # gflat known to be all-zero
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                p.grad = self.gflat[off:off + k].view(p.shape)
                p._cgs_grad = p.grad            # kernels accumulate weight gradients straight into the bucket
                p._cgs_opt = self
                off += k
        self.lr, self.betas, self.eps = lr, betas, eps
        self.group, self.world = process_group, world_size
        if world_size > 1:
            # every rank must start from rank 0's parameters whatever its own seed / checkpoint state was
            torch.distributed.broadcast(self.flat, src=torch.distributed.get_global_rank(process_group, 0)
                                        if process_group is not None else 0, group=process_group)
        # gradient handed over by the whole-step critic kernel as per-CTA partial vectors (ops.critic_train_fused):
        # (buffer, rows, row stride, bucket offset, length), summed in step() instead of RED-accumulated by the kernel
        self.pending_partials = None
        self.adam_done_in_kernel = False
        self._pbuf = None
        self._bar = None
        self._p2p = None
        if world_size > 1 and dev.type == "cuda" and os.environ.get("CGS_P2P", "1") != "0":
            self._p2p = self._setup_p2p(n, dev)

    def _setup_p2p(self, n, dev):
        """Symmetric gradient buffer [2][npad] + flag pad mapped into every peer (one node, NVLink): the all-reduce then
        is `world` peer loads per element inside the Adam kernel (csrc/p2p_adam.cu) instead of an NCCL launch.
        Returns None (-> NCCL all-reduce) if symmetric memory cannot be set up on this system."""
        import ctypes
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            if self.world > 16:
                return None
            npad = (n + 63) // 64 * 64
            sym = symm.empty(2 * npad, dtype=torch.float32, device=dev)
            flags = symm.empty(64, dtype=torch.int32, device=dev)
            # receive buffer of the in-kernel low-latency all-reduce: [2 slots][world][npad] x {value, step tag}
            recv = symm.empty(2 * self.world * npad * 2, dtype=torch.int32, device=dev)
            sym.zero_(); flags.zero_(); recv.zero_()
            torch.cuda.synchronize()
            group = self.group if self.group is not None else dist.group.WORLD
            hs, hf, hr = symm.rendezvous(sym, group), symm.rendezvous(flags, group), symm.rendezvous(recv, group)
            dist.barrier(group=group)
            torch.cuda.synchronize()
            W = self.world
            return dict(npad=npad, sym=sym, flags=flags, recv=recv, hs=hs, hf=hf, hr=hr, rank=dist.get_rank(group),
                        recv_ptrs=(ctypes.c_uint64 * W)(*[int(x) for x in hr.buffer_ptrs]),
                        bufs=(ctypes.c_uint64 * W)(*[int(x) for x in hs.buffer_ptrs]),
                        pads=(ctypes.c_uint64 * W)(*[int(x) for x in hf.buffer_ptrs]),
                        err=torch.zeros(1, dtype=torch.int32, device=dev))
        except Exception as e:          # no peer access / fabric handles on this system: NCCL carries the bucket instead
            print(f"cgs_b200: symmetric-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
            return None

    def check(self):
        """Raise if any bounded device-side wait of this optimizer's kernels ever timed out (grid barrier / peer gradient
        of the whole-step kernel, peer announcement of cgs_p2p_allreduce_adam) or the tcgen05 kernels flagged an mbarrier
        time-out: the kernels skip the affected updates, so the last checkpoint is still good, but training must stop."""
        from . import _lib
        if self.flat.is_cuda and not (self.barrier_ok() and self.p2p_ok() and _lib.lib().cgs_tc_status() == 0
                                      and _lib.lib().cgs_hg_status() == 0 and _lib.lib().cgs_wide_status() == 0):
            raise RuntimeError("cgs_b200: a device-side wait timed out (grid barrier / peer gradient / tcgen05 mbarrier): "
                               "the affected parameter updates were skipped; the optimizer state is not trustworthy")

    def p2p_ok(self):
        """False if a peer ever failed to announce its gradient in time (reads a device flag; synchronises)."""
        return self._p2p is None or int(self._p2p["err"].item()) == 0

    def fused_adam_args(self, off, nparam):
        """cgs_adam_args for the whole-step critic kernel if it may apply this optimizer's update itself: single process,
        and the bucket is exactly the parameter run [off, off + nparam) the kernel owns.  Else None."""
        import ctypes
        from ._lib import AdamArgs
        if off != 0 or nparam != self.flat.numel() or (self.world > 1 and self._p2p is None):
            return None
        if self._bar is None:
            self._bar = torch.zeros(4, dtype=torch.int32, device=self.flat.device)
        a = AdamArgs(self.flat.data_ptr(), self.gflat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.lr,
                     self.betas[0], self.betas[1], self.eps, self.step_count.data_ptr(), self._bar.data_ptr(),
                     1, 0, 0, None)
        if self.world > 1:       # data-parallel: the all-reduce over peer memory happens inside the kernel as well
            q = self._p2p
            a.world, a.rank, a.npad = self.world, q["rank"], q["npad"]
            a.peer_recv = ctypes.cast(q["recv_ptrs"], ctypes.c_void_p)
        return a

    def barrier_ok(self):
        """False if a CTA of the whole-step kernel ever timed out at its grid barrier or waiting for a peer's gradient
        slice (reads a device flag; synchronises)."""
        return self._bar is None or int(self._bar[2].item()) == 0

    def partial_buffer(self, numel):
        if self._pbuf is None or self._pbuf.numel() < numel:
            self._pbuf = torch.empty(numel, device=self.flat.device, dtype=torch.float32)
        return self._pbuf

    def flush_partials(self):
        """Fold a pending partial-gradient hand-over into the bucket (so that `.grad` is complete)."""
        if self.pending_partials is not None:
            buf, rows, stride, off, length = self.pending_partials
            ops.reduce_partials(self.gflat, buf, rows, stride, off, length)
            self.pending_partials = None
            self._clean = False

    def zero_grad(self):
        ops.join_wgrad()
        self.pending_partials = None
        if not self._clean:          # step() already cleared the bucket inside the Adam kernel
            self.gflat.zero_()
            self._clean = True

    def step(self):
        ops.join_wgrad()           # wgrad kernels forked onto the side stream have all landed in the bucket
        if self.adam_done_in_kernel:          # the whole-step critic kernel already applied this update (and cleared the bucket)
            self.adam_done_in_kernel = False
            self._clean = True
            return
        if self._p2p is not None:
            q = self._p2p
            ops.p2p_stage(self.gflat, q["npad"], q["sym"], self.step_count, self.pending_partials)
            self.pending_partials = None
            ops.p2p_allreduce_adam(self.flat, self.m, self.v, q["npad"], q["bufs"], q["pads"], q["rank"], self.world,
                                   self.step_count, q["err"], self.lr, self.betas, self.eps)
            self._clean = True
            return
        if self.world > 1:
            self.flush_partials()
            torch.distributed.all_reduce(self.gflat, group=self.group)   # ranks pre-scale their losses by 1/world
        if self.pending_partials is not None:
            buf, rows, stride, off, length = self.pending_partials
            ops.adam_step_partials(self.flat, self.gflat, self.m, self.v, self.step_count, buf, rows, stride, off, length,
                                   self.lr, self.betas, self.eps)
            self.pending_partials = None
        else:
            ops.adam_step(self.flat, self.gflat, self.m, self.v, self.step_count, self.lr, self.betas, self.eps,
                          clear_grad=True)
        self._clean = True


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def occlude(A, B, Z):
    """`A*(1-Z)+Z*B` on logical-NCHW tensors (reference main.py:395,406), one fused kernel each way."""
    return ops.occlude(_nhwc(A), _nhwc(B), _nhwc(Z)).permute(0, 3, 1, 2)


class Handler:
    def __init__(self, args, device=None, rank=0, world_size=1, process_group=None):
        self.args = args
        argdict = args.__dict__
        self.device = torch.device(device if device is not None else "cuda")
        self.rank, self.world, self.group = rank, world_size, process_group
        self.criticname, self.maskername = "critic", "masker"
        self.reset_models()
        self.models = {self.criticname: self.critic, self.maskername: self.masker}
        # identical naming to reference main.py:86-102
        self.critic_args = "-".join(f"{a}={argdict[a]}" for a in
                                    ["rewidx", "cepochs", "datamode", "datasize", "threshrew", "shift", "chfak", "dropout"]
                                    if argdict[a])
        self.masker_args = "-".join(f"{a}={argdict[a]}" for a in ["mepochs", "L1", "L2", "inject"] if argdict[a])
        self.path = f"{args.name}/"
        self.save_path = self.path + "saves/"
        self.save_paths = {self.criticname: f"{self.save_path}critic-{self.critic_args}.pt",
                           self.maskername: f"{self.save_path}masker-{self.masker_args}.pt"}
        self.contrastive_batchsize = 32      # main.py:309
        self.fused_critic_step = True        # tf32 mode, chfak 1: one kernel per critic_pipe step
        self.critic_bf16 = os.environ.get("CGS_CRITIC_BF16", "1") != "0"   # critic step: bf16 operands (csrc/hg_critic.cu) instead of TF32 (csrc/critic_fused.cu)
        self.wide_critic_step = True         # tensor-core mode, chfak 2..5: TMA-fed tcgen05 convolutions (csrc/wide_tc.cu, wide.py)
        self.hg_score_bf16 = True            # frozen Hourglass step: the three critic scoring passes in ONE bf16 kernel (else TF32)
        self.hg_inference = True             # tensor-core mode, chfak 1: -process in ONE bf16 kernel (csrc/hg_forward.cu)
        self.device_dataset = True           # segmentation_training gathers its batches from a device-resident uint8 dataset
        self.device_dataset_bytes = 8 << 30  # ... when the pos + neg frames fit this budget (100k frames = 1.2 GB)
        self.closs_log, self.seg_log = [], []

    def reset_models(self):
        a = self.args
        self.critic = NewCritic(bottleneck=a.neck, chfak=a.chfak, dropout=a.dropout)      # main.py:108
        self.masker = UnetDecoder(bottleneck=a.neck, chfak=a.chfak)                       # main.py:109
        # dropout Philox keys: a fixed stream index per role (NOT the process-global construction count, so a second
        # reset_models() replays the same stream for the same seed) and the rank, so that data-parallel ranks draw
        # different masks for their shards
        self.critic._instance, self.critic._rank = 1, self.rank
        if a.separate:
            self.sepcrit = NewCritic(bottleneck=a.neck, chfak=a.chfak, dropout=a.dropout)  # main.py:111
            self.sepcrit._instance, self.sepcrit._rank = 2, self.rank
        if self.world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            self._check_shared_seed()

    def _check_shared_seed(self):
        """The DataLoader shuffle, the shift_batch draws (torch RNG) and the contrastive sampling (numpy RNG) must be the
        same on every rank, or the ranks would shard different global batches: require a common torch seed."""
        import torch.distributed as dist
        mine = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64,
                            device=self.device if self.device.type == "cuda" else "cpu")
        seeds = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(seeds, mine, group=self.group)
        if any(int(s) != int(mine) for s in seeds):
            raise RuntimeError("cgs_b200: data-parallel ranks were seeded differently "
                               f"({[int(s) for s in seeds]}): call torch.manual_seed(same) on every rank before Handler()")

    # ------------------------------------------------------------------ data / checkpoints
    def set_data(self, X, Y, I=None, batch_size=64, shuffle=True):
        """`load_data` (main.py:113-129) for arrays already in memory (MineRL collection is out of scope)."""
        a = self.args
        self.X, self.Y = X, Y
        self.I = I if I is not None else np.arange(len(X), dtype=np.int32)
        if a.threshrew:
            self.Y = (self.Y > a.threshrew).astype(float)
        self.train_loader = torch.utils.data.DataLoader(
            torch.utils.data.TensorDataset(torch.from_numpy(self.X), torch.from_numpy(self.Y).t(),
                                           torch.arange(self.X.shape[0], dtype=torch.int32)),
            batch_size=batch_size, shuffle=shuffle)

    def load_models(self, modelnames=()):
        for model in (modelnames or self.models.keys()):
            path = self.save_paths[model]
            if not os.path.exists(path):
                if not self.args.train:
                    print(f"{path} not found")
                return False
            self.models[model].load_state_dict(torch.load(path, map_location=self.device))
        return True

    def save_models(self, modelnames=()):
        if self.rank != 0:
            return
        os.makedirs(self.save_path, exist_ok=True)
        for model in (modelnames or self.models.keys()):
            torch.save(self.models[model].state_dict(), self.save_paths[model])

    # ------------------------------------------------------------------ helpers
    def _opt(self, params):
        return FlatAdam(params, process_group=self.group, world_size=self.world)

    def _shift_roll(self):
        """The two torch.rand(1) draws of shift_batch (main.py:585-586) -> signed roll."""
        xshift = int(self.args.shift * torch.rand(1))
        return xshift if torch.rand(1) > 0.5 else -xshift

    def _to_input(self, X_u8, roll=0):
        """uint8 NHWC (host or device) -> logical NCHW fp32 /255 on device (main.py:189,360)."""
        x = X_u8 if torch.is_tensor(X_u8) else torch.from_numpy(np.ascontiguousarray(X_u8))
        x = x.to(self.device, non_blocking=True)
        if torch.is_tensor(roll):                      # device int32 scalar: graph-replayable shift
            return ops.frames_to_float(x, 0, roll).permute(0, 3, 1, 2)
        return ops.frames_to_float(x, roll).permute(0, 3, 1, 2)

    def _shard(self, n):
        """This rank's slice of a global batch of n: balanced (sizes differ by at most one) and never empty.  A batch
        smaller than the world size is processed whole by every rank (the mean over ranks of identical gradients is the
        global-batch gradient), so no rank ever sits out a collective."""
        if n < self.world:
            return slice(0, n)
        q, r = divmod(n, self.world)
        lo = self.rank * q + min(self.rank, r)
        return slice(lo, lo + q + (1 if self.rank < r else 0))

    def _shard_weight(self, n):
        """d(global mean loss) / d(this rank's local mean loss) = B_local / B_global (1/world for a replicated batch):
        the gradient sum over ranks is then the reference's global-batch mean gradient even for ragged shards."""
        if self.world == 1:
            return 1.0
        if n < self.world:
            return 1.0 / self.world
        sl = self._shard(n)
        return (sl.stop - sl.start) / n

    # ------------------------------------------------------------------ critic regression
    def critic_step(self, X_u8, Y, opti, roll=0, weight=None):
        """Loop body of critic_pipe (main.py:185-200) on this rank's shard; returns the loss tensor.  `weight` =
        B_local / B_global (default 1/world: equal shards)."""
        a = self.args
        weight = (1.0 / self.world) if weight is None else float(weight)
        x = X_u8 if torch.is_tensor(X_u8) else torch.from_numpy(np.ascontiguousarray(X_u8))
        x = x.to(self.device, non_blocking=True)
        Yd = Y.to(self.device, non_blocking=True).float()
        if self.fused_critic_step and ops.critic_fused_supported(self.critic) and x.dtype == torch.uint8:
            # whole step in one kernel: every activation of a frame stays in shared memory (csrc/critic_fused.cu)
            rng = self.critic._dropout_rng(x.device)               # masks drawn in the kernel (same Philox stream) ...
            masks = (None, None, None) if rng is not None else self.critic._dropout_masks(x.shape[0], x.device)   # ... or forced / none
            opti.zero_grad()
            loss, _ = ops.critic_train_fused(self.critic, x.contiguous(), Yd.contiguous(), roll, masks,
                                             loss_grad=weight, bce=bool(a.threshrew), rng=rng,
                                             fuse_adam=bool(getattr(opti, "_clean", False)),
                                             bf16=self.critic_bf16 and isinstance(opti, FlatAdam))
            opti.step()
            return loss
        if self.wide_critic_step and x.dtype == torch.uint8 and wide.supported(self.critic):
            # chfak 2..5: bf16 chunk-planar activations, TMA-fed tcgen05 forward / input / weight gradients (wide.py)
            rng = self.critic._dropout_rng(x.device)
            masks = (None, None, None) if rng is not None else self.critic._dropout_masks(x.shape[0], x.device)
            opti.zero_grad()
            loss, _ = wide.critic_train_wide(self.critic, x.contiguous(), Yd.contiguous(), roll, masks,
                                             loss_grad=weight, bce=bool(a.threshrew), rng=rng)
            opti.step()
            return loss
        pred = self.critic.forward_frames(x, roll).squeeze(1)     # cast + roll fused into features.0's operand load
        loss = ops.pred_loss(pred, Yd, bce=bool(a.threshrew))
        opti.zero_grad()
        (loss * weight if self.world > 1 else loss).backward()
        opti.step()
        return loss.detach()

    def critic_pipe(self, mode="train"):
        a = self.args
        if a.cload and self.load_models([self.criticname]):
            print("loaded critic, no new training")
            return
        critic = self.critic.to(self.device)
        opti = self._opt(critic.parameters())
        for epoch in range(int(mode == "test") or a.cepochs):
            for b_idx, (X, Y, I) in enumerate(self.train_loader):
                roll = self._shift_roll() if a.shift else 0
                sl = self._shard(len(X))
                loss = self.critic_step(X[sl], Y[sl, a.rewidx], opti, roll, weight=self._shard_weight(len(X)))
                self.closs_log.append(loss)
            if not (epoch + 1) % a.saveevery:
                opti.check()                     # never checkpoint past a skipped update
                self.save_models([self.criticname])
        self.closs_log = [float(v) for v in torch.stack(self.closs_log).cpu()] if self.closs_log else []
        opti.check()

    # ------------------------------------------------------------------ pos / neg split
    def extract_contrastive_data(self):
        a = self.args
        critic = self.critic.to(self.device).eval()
        batchsize = 128
        preds = []
        with torch.no_grad():
            for bidx in range(math.ceil(len(self.X) / batchsize)):
                chunk = self.X[bidx * batchsize:(bidx + 1) * batchsize]
                if self.fused_critic_step and ops.critic_fused_supported(critic):      # raw uint8 frames in, ONE kernel
                    xu8 = torch.from_numpy(np.ascontiguousarray(chunk)).to(self.device, non_blocking=True)
                    preds.append(ops.critic_forward_frames(critic, xu8).squeeze(1))
                else:
                    preds.append(critic(self._to_input(chunk)).squeeze(1))
        preds = torch.cat(preds, dim=0).cpu()
        positives = (preds > a.high_rew_thresh).numpy()
        negatives = (preds < a.low_rew_thresh).numpy()
        assert positives.sum() >= 500 and negatives.sum() >= 500          # main.py:281
        self.Xpos, self.Ypos = self.X[positives], self.Y[:, positives]
        self.Xneg, self.Yneg = self.X[negatives], self.Y[:, negatives]
        assert preds[torch.from_numpy(positives)].mean() > a.high_rew_thresh   # main.py:302
        self.XposIdxs = np.arange(len(self.Xpos))
        self.XnegIdxs = np.arange(len(self.Xneg))
        self.ContrastIdxs = np.arange(len(self.Xneg))
        cb = self.contrastive_batchsize
        self.get_contrastive_idxs = lambda: (np.random.choice(self.XposIdxs, cb),
                                             np.random.choice(self.XnegIdxs, cb),
                                             np.random.choice(self.ContrastIdxs, 2 * cb))
        self.preds = preds

    # ------------------------------------------------------------------ Hourglass training
    def segmentation_losses(self, A, B, Y=None):
        """Loss terms of one segmentation_training step (main.py:364-429).  A, B logical NCHW."""
        a = self.args
        critic, masker = self.critic, self.masker
        pred, embeds = critic(A, collect=True)
        with torch.no_grad():
            if self.fused_critic_step and ops.critic_fused_supported(critic):      # forward only, one kernel
                rng = critic._dropout_rng(B.device)
                masks = (None, None, None) if rng is not None else critic._dropout_masks(B.shape[0], B.device)
                negpred = ops.critic_forward_fused(critic, B.permute(0, 2, 3, 1), masks, rng).squeeze(1)
            else:
                negpred = critic(B).squeeze(1)
        pred = pred.squeeze(1)
        terms = {}
        loss = 0
        if a.live:
            terms["critic"] = ops.pred_loss(pred, Y, bce=bool(a.threshrew))
            loss = loss + a.lfak * terms["critic"]
        if a.separate:
            _, embeds = self.sepcrit(A, collect=True)
        Z = masker(A, embeds)
        replaced = occlude(A, B, Z)
        # frozen critic (tf32 mode, chfak 1): critic(blend) + loss + the backward into the blend in ONE kernel
        frozen = self.fused_critic_step and ops.critic_fused_supported(critic) and not any(q.requires_grad for q in critic.parameters())

        def scored(blend, target):
            if frozen:
                rng = critic._dropout_rng(blend.device)
                masks = (None, None, None) if rng is not None else critic._dropout_masks(blend.shape[0], blend.device)
                return ops.critic_loss_xgrad(critic, blend.permute(0, 2, 3, 1), target, masks, rng)
            return ops.pred_loss(critic(blend).squeeze(1), target)
        terms["replace"] = scored(replaced, negpred)
        loss = loss + terms["replace"]
        if a.inject:
            injected = occlude(B, A, Z)
            terms["inject"] = scored(injected, pred.detach())
            loss = loss + terms["inject"]
        vpred = None if a.staticnorm else pred.detach()
        if a.L1:
            terms["L1"] = ops.mask_reg(_nhwc(Z), vpred, l1=a.L1)
            loss = loss + terms["L1"]
        if a.L2:
            terms["L2"] = ops.mask_reg(_nhwc(Z), vpred, l2=a.L2)
            loss = loss + terms["L2"]
        return loss, terms, Z

    def _hg_fused(self, opti):
        """True when one frozen-critic Hourglass step can run as the whole-frame kernels (csrc/hg_*.cu + cgs_hg_score):
        chfak-1 geometry in tensor-core mode, critic out of the optimizer, and the optimizer's bucket is exactly the
        masker's parameters in state_dict order."""
        a = self.args
        if a.live or a.separate or not self.fused_critic_step or not isinstance(opti, FlatAdam):
            return False
        if not (ops.hg_supported(self.critic, self.masker) and ops.critic_fused_supported(self.critic)):
            return False
        mp = list(self.masker.parameters())
        return len(opti.params) == len(mp) and all(q is r for q, r in zip(opti.params, mp))

    def segmentation_step_fused(self, X_u8, CX_u8, opti, roll=0, weight=1.0):
        """One frozen-critic iteration of segmentation_training (main.py:344-463) in five launches: weight-fragment pack,
        Hourglass forward on A (critic with embeds + decoder + masker; leaves the tape), the critic scoring passes (critic(B),
        the two scored blends with their losses, regulariser and d loss / d mask: cgs_hg_score_bf16), the masker's whole
        backward, and the partial-vector sum + Adam.  No activation but the mask, its gradient and the 54 KB/frame bf16 tape touches HBM."""
        a = self.args
        critic, masker, dev = self.critic, self.masker, self.device
        x = X_u8 if torch.is_tensor(X_u8) else torch.from_numpy(np.ascontiguousarray(X_u8))
        cx = CX_u8 if torch.is_tensor(CX_u8) else torch.from_numpy(np.ascontiguousarray(CX_u8))
        x, cx = x.to(dev, non_blocking=True).contiguous(), cx.to(dev, non_blocking=True).contiguous()
        B = x.shape[0]
        st = getattr(self, "_hg_state", None)
        if st is None or st["B"] != B or st["tape"].device != x.device:
            L = _lib.lib()
            st = self._hg_state = dict(B=B, tape=ops.hg_tape(B, x.device),
                                       pack=torch.empty(L.cgs_hg_pack_words(), device=x.device, dtype=torch.int32))
        pack = ops.hg_pack(critic, masker, out=st["pack"])

        def drop():        # the critic's dropout for its next forward pass, in the reference's call order
            rng = critic._dropout_rng(dev)
            return (rng, None) if rng is not None else (None, critic._dropout_masks(B, dev))
        rng, masks = drop()                                                                # critic(A, collect=True)   main.py:364
        pred, Z, _ = ops.hg_forward(critic, masker, x, roll=roll, train=critic.training, masks=masks, rng=rng, tape=st["tape"],
                                    pack=pack)
        vpred = None if a.staticnorm else pred.squeeze(1)
        if self.hg_score_bf16:
            # critic(B), critic(replaced), critic(injected) with their losses, the regulariser and d/dZ: ONE bf16 kernel
            rng = critic._dropout_rng(dev)
            forced = None
            if rng is None:
                forced = [critic._dropout_masks(B, dev) for _ in range(3 if a.inject else 2)] + ([None] if not a.inject else [])
            losses, dz, _, _, _ = ops.hg_score_bf16(critic, x, cx, Z, pack, None, pred.squeeze(1) if a.inject else None, roll=roll,
                                                    masks=forced, rng=rng, loss_grad=weight, vpred=vpred,
                                                    l1=float(a.L1 or 0.0), l2=float(a.L2 or 0.0))
        else:
            rng, masks = drop()                                                            # critic(B)                 main.py:365
            negpred = ops.critic_forward_frames(critic, cx, 0, masks if masks is not None else (None, None, None), rng)
            if rng is not None:
                m_r = m_i = None
            else:
                m_r = critic._dropout_masks(B, dev)                                        # critic(replaced)          main.py:396
                m_i = critic._dropout_masks(B, dev) if a.inject else None                  # critic(injected)          main.py:407
                m_r = None if m_r[0] is None else m_r
                m_i = None if (m_i is None or m_i[0] is None) else m_i
            losses, dz, _, _ = ops.hg_score(critic, x, cx, Z, negpred.squeeze(1), pred.squeeze(1) if a.inject else None, roll=roll,
                                            masks=m_r, masks_inject=m_i, rng=critic._dropout_rng(dev) if rng is not None else None,
                                            loss_grad=weight, vpred=vpred, l1=float(a.L1 or 0.0), l2=float(a.L2 or 0.0))
        opti.zero_grad()
        L = _lib.lib()
        grid, stride = L.cgs_hg_grid(B), L.cgs_hg_partial_stride()
        buf = opti.partial_buffer(grid * stride)
        ops.hg_backward(masker, x, st["tape"], Z, dz, roll=roll, pack=pack, partials=buf)
        opti.pending_partials = (buf, grid, stride, 0, opti.flat.numel())
        opti.step()
        terms = {"replace": losses[0]}
        if a.inject:
            terms["inject"] = losses[1]
        if a.L1:
            terms["L1"] = losses[2]
        if a.L2:
            terms["L2"] = losses[3]
        self._last_mask = Z
        return terms

    def segmentation_step(self, X_u8, CX_u8, Y, opti, roll=0, weight=None):
        weight = (1.0 / self.world) if weight is None else float(weight)
        if self._hg_fused(opti) and (torch.is_tensor(X_u8) and X_u8.dtype == torch.uint8 or not torch.is_tensor(X_u8)):
            return self.segmentation_step_fused(X_u8, CX_u8, opti, roll, weight if self.world > 1 else 1.0)
        A = self._to_input(X_u8, roll)
        B = self._to_input(CX_u8)
        Yd = None if Y is None else Y.to(self.device).float()
        loss, terms, _ = self.segmentation_losses(A, B, Yd)
        opti.zero_grad()
        (loss * weight if self.world > 1 else loss).backward()
        opti.step()
        return {k: v.detach() for k, v in terms.items()}

    def segmentation_training(self):
        a = self.args
        self.extract_contrastive_data()
        critic = self.critic.to(self.device).train()
        masker = self.masker.to(self.device).train()
        sep = [self.sepcrit.to(self.device).train()] if a.separate else []
        if a.live:
            opti = self._opt(chain(critic.parameters(), masker.parameters(), *[s.parameters() for s in sep]))
        else:
            # the reference leaves requires_grad on and merely omits the critic from Adam (main.py:334);
            # its critic gradients are never read, so they are not computed here
            for p in critic.parameters():
                p.requires_grad_(False)
            opti = self._opt(chain(masker.parameters(), *[s.parameters() for s in sep]))
        # device-resident dataset: the pos / neg frames are uploaded ONCE as uint8 and every step's
        # `Xpos[Hidx]`, `Xneg[Lidx]`, `Xneg[Cidx]` (main.py:345-353) is an index gather on the device (cgs_gather_frames)
        # fed by a few hundred bytes of indices instead of the frames themselves
        data = None
        if self.device_dataset and self.device.type == "cuda" and self.Xpos.nbytes + self.Xneg.nbytes <= self.device_dataset_bytes:
            data = torch.from_numpy(np.ascontiguousarray(np.concatenate((self.Xpos, self.Xneg), axis=0))).to(self.device)
            labels = torch.from_numpy(np.concatenate((self.Ypos[a.rewidx], self.Yneg[a.rewidx]))).to(self.device)
            npos = len(self.Xpos)
        try:
            for epoch in range(a.mepochs):
                for b_idx in range(math.ceil(self.Xpos.shape[0] / self.contrastive_batchsize)):
                    Hidx, Lidx, Cidx = self.get_contrastive_idxs()
                    roll = self._shift_roll() if a.shift else 0
                    n = len(Hidx) + len(Lidx)
                    sl = self._shard(n)
                    if data is not None:
                        xi = np.concatenate((Hidx, npos + Lidx))[sl]
                        ci = (npos + Cidx)[sl]
                        idx = torch.from_numpy(np.concatenate((xi, ci)).astype(np.int32)).to(self.device, non_blocking=True)
                        frames = ops.gather_frames(data, idx)
                        X, CX = frames[:len(xi)], frames[len(xi):]
                        Y = labels[idx[:len(xi)].long()]
                    else:
                        X = np.concatenate((self.Xpos[Hidx], self.Xneg[Lidx]), axis=0)[sl]
                        Y = torch.from_numpy(np.concatenate((self.Ypos[a.rewidx, Hidx], self.Yneg[a.rewidx, Lidx])))[sl]
                        CX = self.Xneg[Cidx][sl]
                    self.seg_log.append(self.segmentation_step(X, CX, Y, opti, roll, weight=self._shard_weight(n)))
                if not (epoch + 1) % a.saveevery:
                    opti.check()
                    self.save_models([self.maskername])
            opti.check()
        finally:
            if not a.live:
                for p in critic.parameters():
                    p.requires_grad_(True)
        self.seg_log = [{k: float(v) for k, v in t.items()} for t in self.seg_log]

    # ------------------------------------------------------------------ -process
    def segment_device(self, xu8, threshold=None):
        """Loop body of Handler.segment (main.py:1139-1164) for one batch of uint8 NHWC frames already on the device:
        (pred [B,1], mask [B,1,64,64], hard uint8 [B,1,64,64] or None), all device tensors.  Modules must already be in
        the wanted train/eval mode.  No autograd."""
        a = self.args
        critic, masker = self.critic, self.masker
        with torch.no_grad():
            if not a.separate and self.hg_inference and ops.hg_supported(critic, masker) and not critic.training:
                return ops.hg_forward(critic, masker, xu8, thresh=threshold or None)      # ONE kernel, bf16 operands
            if not a.separate and ops.infer_fused_supported(critic, masker):
                pred, o0 = ops.infer_encode_decode(critic, masker, xu8)
                mask, hm = ops.masker_fused(masker, xu8, o0, threshold or None)
                return pred, mask, hm
            batch = ops.frames_to_float(xu8, 0).permute(0, 3, 1, 2)
            pred, embeds = critic(batch, collect=True)
            if a.separate:
                _, embeds = self.sepcrit.to(self.device).train(critic.training)(batch, collect=True)
            if threshold:
                mask, hm = masker.forward_hard(batch, embeds, threshold)
                return pred, mask, hm
            return pred, masker(batch, embeds), None

    def segment_arrays(self, X_u8, batchsize=128):
        """Loop of Handler.segment (main.py:1130-1167) on uint8 frames: returns (preds, M, hardM)."""
        a = self.args
        train = bool(a.noevalmode)
        self.critic.to(self.device).train(train)
        self.masker.to(self.device).train(train)
        preds, M, hard = [], [], []
        for bidx in range(0, len(X_u8), batchsize):
            xu8 = torch.from_numpy(np.ascontiguousarray(X_u8[bidx:bidx + batchsize])).to(self.device)
            pred, mask, hm = self.segment_device(xu8, a.binarymaskthreshold)
            if hm is not None:
                hard.append(hm.cpu().numpy().astype(bool))
            preds.append(pred.squeeze(1).cpu().numpy())
            M.append(mask.cpu().numpy())
        M = np.concatenate(M, axis=0)
        return np.concatenate(preds, axis=0), M, (np.concatenate(hard, axis=0) if hard else None)

    def eval_iou(self, X_u8, GT, batchsize=128):
        """The IoU of Handler.eval (reference main.py:891-1017 without CRF / saliency / plots): masks of all frames,
        `hardM = M > eval_thresh` (main.py:964), `get_iou(hardM, Y)` (main.py:1005, 1265-1270).  GT: bool/uint8 [N,64,64]
        ground-truth masks.  Compare + intersection/union counts run on the device; returns (iou rounded as the reference
        does, intersection, union)."""
        a = self.args
        train = bool(a.noevalmode)
        critic = self.critic.to(self.device).train(train)
        masker = self.masker.to(self.device).train(train)
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        GT = np.ascontiguousarray(np.asarray(GT).astype(np.uint8))
        with torch.no_grad():
            for bidx in range(0, len(X_u8), batchsize):
                xu8 = torch.from_numpy(np.ascontiguousarray(X_u8[bidx:bidx + batchsize])).to(self.device)
                gt = torch.from_numpy(GT[bidx:bidx + batchsize]).to(self.device)
                _, mask, _ = self.segment_device(xu8, None)
                ops.iou_counts(mask, gt, a.eval_thresh, counts, strict=True)
        inter, union = (int(v) for v in counts.cpu())
        return (round(inter / union, 3) if union else float("nan")), inter, union

    def eval_saliency(self, X_u8, GT=None, batchsize=128):
        """The saliency baseline of Handler.eval / Handler.segment (reference main.py:941-951, 974-993, 1010-1011): per batch
        `pred.mean().backward(); batch.grad.abs().sum(1)` (ONE kernel: critic forward + input gradient), then over ALL frames the
        normalisation (`-salglobal`: mean norm; else the per-frame k-th value), `* pred`, clip, `> salience_thresh`, and, given
        ground-truth masks, `get_iou(salhardM, Y)`.  Everything stays on the device.
        Returns (salM [N,1,64,64], salhardM uint8 [N,1,64,64], iou or None)."""
        a = self.args
        train = bool(a.noevalmode)
        critic = self.critic.to(self.device).train(train)
        sal, preds = [], []
        for bidx in range(0, len(X_u8), batchsize):
            xu8 = torch.from_numpy(np.ascontiguousarray(X_u8[bidx:bidx + batchsize])).to(self.device)
            pred, m = ops.critic_saliency(critic, ops.frames_to_float(xu8, 0))     # frames / 255 (main.py:921, 1127)
            sal.append(m); preds.append(pred)
        sal, preds = torch.cat(sal), torch.cat(preds)
        salM, salhard = ops.saliency_normalize(sal, preds, a.salience_thresh, global_norm=bool(a.salglobal))
        iou = None
        if GT is not None:
            counts = torch.zeros(2, dtype=torch.int64, device=self.device)
            gt = torch.from_numpy(np.ascontiguousarray(np.asarray(GT).astype(np.uint8))).to(self.device)
            ops.iou_counts(salhard.float(), gt, 0.5, counts, strict=True)
            inter, union = (int(v) for v in counts.cpu())
            iou = round(inter / union, 3) if union else float("nan")
        return salM, salhard, iou

    def segment(self, folder):
        """`-process`: PNG folder in, mask PNGs out (main.py:1103-1223): `raw-mask` and `thresholded-mask` per frame, or one
        `-concatenated` strip frame | raw | thresholded.  The uint8 x 3-channel image rows come straight from the device
        (cgs_mask_images); the host only encodes PNG."""
        from PIL import Image
        a = self.args
        names = os.listdir(folder)
        X = np.stack([np.array(Image.open(f"{folder}/{n}")) for n in names]).astype(np.uint8)
        names = [n[:-1 - n[::-1].index(".")] for n in names if "." in n]
        train = bool(a.noevalmode)
        self.critic.to(self.device).train(train)
        self.masker.to(self.device).train(train)
        out = a.mask_output_imgs
        os.makedirs(out, exist_ok=True)
        preds, M, hard = [], [], []
        for bidx in range(0, len(X), 128):
            xu8 = torch.from_numpy(np.ascontiguousarray(X[bidx:bidx + 128])).to(self.device)
            pred, mask, hm = self.segment_device(xu8, a.binarymaskthreshold)
            if hm is None:                                   # threshold 0: every pixel passes `M >= 0`
                hm = torch.ones_like(mask, dtype=torch.uint8)
            if a.concatenated:
                strips = ops.mask_images(mask, hm, xu8, concatenated=True).cpu().numpy()
                for i in range(len(strips)):
                    Image.fromarray(strips[i]).save(f"{out}/{names[bidx + i]}_with_mask.png")
            else:
                raw, thr = (t.cpu().numpy() for t in ops.mask_images(mask, hm))
                for i in range(len(raw)):
                    Image.fromarray(raw[i]).save(f"{out}/{names[bidx + i]}-raw-mask.png")
                    Image.fromarray(thr[i]).save(f"{out}/{names[bidx + i]}-thresholded-mask.png")
            preds.append(pred.squeeze(1).cpu().numpy()); M.append(mask.cpu().numpy()); hard.append(hm.cpu().numpy().astype(bool))
        return np.concatenate(preds), np.concatenate(M), np.concatenate(hard)
