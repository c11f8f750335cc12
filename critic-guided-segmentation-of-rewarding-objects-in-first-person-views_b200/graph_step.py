"""CUDA-graph captured training / inference steps.

The reference loops are host-bound (one Python-dispatched ATen kernel per op and 3-4
`.item()` syncs per step); here one whole step — uint8->float + shift roll,
forward, loss, backward, [gradient all-reduce], Adam — is captured once on static buffers
and replayed with a single launch.  Public API:

    step = GraphedCriticStep(handler, batch=256)        # or GraphedHourglassStep / GraphedSegment
    loss = step(X_u8_host_pinned, Y_host_pinned)        # H2D copy, replay, returns device scalar
"""
import torch

from . import ops
from .train_handler import FlatAdam


def _train_state(opti, *critics):
    """Every device tensor a training step mutates besides its inputs: parameters, gradient bucket, Adam moments and step
    counter, the dropout Philox call counters of the critics."""
    st = [opti.flat, opti.gflat, opti.m, opti.v, opti.step_count[:2]]     # NOT the exchange epoch [2]: peers' flags only grow
    for c in critics:
        if c is not None:
            c._dropout_rng(opti.flat.device)          # creates the counter if this module has not drawn yet
            if c._rng_state is not None:
                st.append(c._rng_state)
    return st


def _capture(fn, warmup=3, state=()):
    """Standard whole-step capture: warm up on a side stream, then capture `fn` into a graph.  The warm-up runs REAL steps
    (optimizer update, RNG advance) on whatever the static buffers hold, so every tensor in `state` is snapshotted before
    and restored after: constructing a graphed step leaves parameters, Adam moments, step count and dropout stream exactly
    as they were (the reference trajectory starts at the first replay)."""
    torch.cuda.synchronize()
    saved = [t.clone() for t in state]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warmup):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = ops.launch_count()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        out = fn()
    n = ops.launch_count() - n0
    for t, v in zip(state, saved):
        t.copy_(v)
    if saved:
        ops.weights_changed()
    torch.cuda.synchronize()
    return g, out, n


class _Graphed:
    launches = 0
    trains = False          # replaying rewrites parameters (optimizer kernels inside the graph)

    def load(self, *host):
        for dst, src in zip(self.static_in, host):
            dst.copy_(src, non_blocking=True)

    def replay(self):
        self.graph.replay()
        if self.trains:
            ops.weights_changed()
        return self.out

    def __call__(self, *host):
        self.load(*host)
        return self.replay()


class GraphedCriticStep(_Graphed):
    """One critic_pipe iteration (reference main.py:185-200) as a CUDA graph.  The shift_batch roll is a
    device int32 (`self.roll`), so it can change per replay: `step.roll.fill_(r)` before the call."""
    trains = True

    def __init__(self, handler, batch, opti=None, X=None, Y=None, warmup=3):
        H = self.H = handler
        dev = H.device
        H.critic.to(dev).train()
        self.opti = opti or FlatAdam(H.critic.parameters(), process_group=H.group, world_size=H.world)
        # static inputs: own buffers, or caller-provided device tensors (e.g. one slice of a resident dataset per graph)
        self.X = X if X is not None else torch.zeros((batch, 64, 64, 3), dtype=torch.uint8, device=dev)
        self.Y = Y if Y is not None else torch.zeros((batch,), dtype=torch.float32, device=dev)
        self.roll = torch.zeros(1, dtype=torch.int32, device=dev)
        self.static_in = (self.X, self.Y)

        def fn():
            return H.critic_step(self.X, self.Y, self.opti, roll=self.roll)
        self.graph, self.out, self.launches = _capture(fn, warmup=warmup, state=_train_state(self.opti, H.critic))


class GraphedHourglassStep(_Graphed):
    """One segmentation_training iteration (reference main.py:344-463) as a CUDA graph."""
    trains = True

    def __init__(self, handler, batch, opti=None, X=None, CX=None, Y=None, warmup=3):
        H = self.H = handler
        dev = H.device
        a = H.args
        H.critic.to(dev).train()
        H.masker.to(dev).train()
        if a.live:
            params = list(H.critic.parameters()) + list(H.masker.parameters())
        else:
            for p in H.critic.parameters():
                p.requires_grad_(False)
            params = list(H.masker.parameters())
        self.opti = opti or FlatAdam(params, process_group=H.group, world_size=H.world)
        # static inputs: own buffers, or caller-provided device tensors (one slice of a resident dataset per graph)
        self.X = X if X is not None else torch.zeros((batch, 64, 64, 3), dtype=torch.uint8, device=dev)
        self.CX = CX if CX is not None else torch.zeros((batch, 64, 64, 3), dtype=torch.uint8, device=dev)
        self.Y = Y if Y is not None else torch.zeros((batch,), dtype=torch.float32, device=dev)
        self.roll = torch.zeros(1, dtype=torch.int32, device=dev)
        self.static_in = (self.X, self.CX, self.Y)

        def fn():
            terms = H.segmentation_step(self.X, self.CX, self.Y, self.opti, roll=self.roll)
            return torch.stack([terms[k] for k in sorted(terms)])
        self.graph, self.out, self.launches = _capture(fn, warmup=warmup, state=_train_state(self.opti, H.critic,
                                                                              getattr(H, "sepcrit", None) if a.separate else None))
        self.term_names = sorted(k for k in ("critic", "replace", "inject", "L1", "L2")
                                 if (k != "critic" or a.live) and (k != "inject" or a.inject)
                                 and (k != "L1" or a.L1) and (k != "L2" or a.L2))


class GraphedSegment(_Graphed):
    """One batch of Handler.segment (reference main.py:1134-1164): critic(collect) -> masker -> >= threshold.
    The captured graph reads the decoder weights through the fragment buffer packed at capture time: build a new
    GraphedSegment after the masker has been trained further."""

    def __init__(self, handler, batch, threshold=0.1, X=None):
        H = self.H = handler
        dev = H.device
        H.critic.to(dev).eval()
        H.masker.to(dev).eval()
        self.X = X if X is not None else torch.zeros((batch, 64, 64, 3), dtype=torch.uint8, device=dev)
        self.static_in = (self.X,)

        def fn():
            return H.segment_device(self.X, threshold)       # whole-frame kernels when the geometry allows, else per layer
        self.graph, self.out, self.launches = _capture(fn)


class HostPipeline:
    """End-to-end feed for any graphed step from HOST batches: `slots` are >= 2 graphed steps with their own static input
    buffers (training steps share one model and one optimizer).  The H2D copy of batch i+1 runs on a copy stream while
    slot i replays; `result_of(slot)` (a device tensor: loss terms, hard masks ...) is read back after every step with an
    asynchronous D2H copy into a pinned ring of `ring` entries.  `step()` never blocks the host; `results()` synchronises."""

    def __init__(self, slots, result_of, ring=8):
        self.slots, self.result_of = slots, result_of
        self.copy_stream = torch.cuda.Stream()
        self.ready = [torch.cuda.Event() for _ in slots]
        self.done = [torch.cuda.Event() for _ in slots]
        r0 = result_of(slots[0])
        self.ring = torch.zeros((ring,) + tuple(r0.shape), dtype=r0.dtype).pin_memory()
        self.d2h_bytes = r0.numel() * r0.element_size()
        self.i = 0

    def step(self, *host):
        k = self.i % len(self.slots)
        st = self.slots[k]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.done[k])          # the replay that last read this slot's inputs has finished
            st.load(*host)
            self.ready[k].record(self.copy_stream)
        main = torch.cuda.current_stream()
        main.wait_event(self.ready[k])
        st.replay()
        self.ring[self.i % self.ring.shape[0]].copy_(self.result_of(st), non_blocking=True)
        self.done[k].record(main)
        self.i += 1

    def results(self):
        torch.cuda.synchronize()
        return self.ring


class PipelinedCriticTrainer:
    """End-to-end critic training from HOST batches: pinned uint8 frames + labels in, loss values out.

    Two captured step graphs with their own static input buffers share one model and one FlatAdam; the H2D
    copy of batch i+1 runs on a copy stream while graph i replays, and every step's loss is read back with an
    asynchronous D2H copy into a pinned ring (the reference syncs 1-4 times per step with `.item()`,
    main.py:196-200).  `step()` never blocks the host; `losses()` synchronises and returns the ring."""

    def __init__(self, handler, batch, depth=2, ring=4096):
        self.opti = FlatAdam(handler.critic.to(handler.device).parameters(), process_group=handler.group,
                             world_size=handler.world)
        self.handler, self.batch = handler, batch
        self._cslots = None
        self.slots = [GraphedCriticStep(handler, batch, self.opti) for _ in range(depth)]
        self.launches = self.slots[0].launches
        self.copy_stream = torch.cuda.Stream()
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.loss_ring = torch.zeros(ring, dtype=torch.float32).pin_memory()
        self.i = 0

    def step(self, X_host, Y_host, roll=None):
        k = self.i % len(self.slots)
        st = self.slots[k]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.done[k])          # the graph that last read this slot has finished
            st.X.copy_(X_host, non_blocking=True)
            st.Y.copy_(Y_host, non_blocking=True)
            if roll is not None:
                st.roll.fill_(int(roll))
            self.ready[k].record(self.copy_stream)
        main = torch.cuda.current_stream()
        main.wait_event(self.ready[k])
        st.graph.replay()
        ops.weights_changed()
        self.done[k].record(main)
        self.loss_ring[self.i % self.loss_ring.numel()].copy_(st.out, non_blocking=True)
        self.i += 1

    def _chunk_slot(self, chunk):
        """Static buffers for `chunk` consecutive batches + ONE captured graph that runs the `chunk` steps on them."""
        H, B = self.handler, self.batch
        dev = H.device
        Xd = torch.zeros((chunk * B, 64, 64, 3), dtype=torch.uint8, device=dev)
        Yd = torch.zeros(chunk * B, dtype=torch.float32, device=dev)
        rolls = torch.zeros(chunk, dtype=torch.int32, device=dev)

        def fn():
            return torch.stack([H.critic_step(Xd[k * B:(k + 1) * B], Yd[k * B:(k + 1) * B], self.opti, roll=rolls[k:k + 1])
                                for k in range(chunk)])
        graph, out, _ = _capture(fn, warmup=1, state=_train_state(self.opti, H.critic))
        return dict(X=Xd, Y=Yd, rolls=rolls, graph=graph, out=out, ready=torch.cuda.Event(), done=torch.cuda.Event())

    def train(self, X_host, Y_host, rolls=None, chunk=None):
        """Run len(X_host) // batch steps over a pinned host dataset (critic_pipe's inner loop, main.py:185-200).
        Work is issued per CHUNK of `chunk` batches: one host->device copy (keeps PCIe at its streaming rate; a 3 MB copy
        per step does not), ONE graph launch that runs the `chunk` steps straight from the chunk buffer, one read-back of
        the `chunk` losses; two chunk buffers alternate so the copy of chunk i+1 overlaps the steps of chunk i.
        A ragged tail goes through step().  Never blocks the host."""
        B = self.batch
        n = X_host.shape[0] // B
        if chunk is None:
            # measured on B200 / PCIe Gen5: a 25 MB pinned copy streams at 31 GB/s, a 50 MB one at 54 GB/s (tools/e2e_probe.py)
            chunk = max(1, min(64, -(-48 * 2 ** 20 // (B * 12288))))
        if self._cslots is None or self._cslots[0]["rolls"].numel() != chunk:
            self._cslots = [self._chunk_slot(chunk) for _ in range(2)]
            self._ci = 0
        main = torch.cuda.current_stream()
        c = 0
        while n - c >= chunk:
            sl = self._cslots[self._ci % 2]
            self._ci += 1
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(sl["done"])            # the steps that last read this chunk buffer are finished
                sl["X"].copy_(X_host[c * B:(c + chunk) * B], non_blocking=True)
                sl["Y"].copy_(Y_host[c * B:(c + chunk) * B], non_blocking=True)
                if rolls is not None:
                    sl["rolls"].copy_(torch.as_tensor(rolls[c:c + chunk], dtype=torch.int32), non_blocking=True)
                sl["ready"].record(self.copy_stream)
            main.wait_event(sl["ready"])
            sl["graph"].replay()
            ops.weights_changed()
            sl["done"].record(main)
            r = self.i % self.loss_ring.numel()
            head = min(chunk, self.loss_ring.numel() - r)       # the ring wraps: split the read-back, never drop a loss
            self.loss_ring[r:r + head].copy_(sl["out"][:head], non_blocking=True)
            if head < chunk:
                self.loss_ring[:chunk - head].copy_(sl["out"][head:], non_blocking=True)
            self.i += chunk
            c += chunk
        for k in range(c, n):
            self.step(X_host[k * B:(k + 1) * B], Y_host[k * B:(k + 1) * B], None if rolls is None else rolls[k])
        return n

    def losses(self):
        """The most recent min(steps, ring) losses in step order (synchronises)."""
        torch.cuda.synchronize()
        self.opti.check()
        R = self.loss_ring.numel()
        if self.i <= R:
            return self.loss_ring[:self.i].clone()
        r = self.i % R
        return torch.cat((self.loss_ring[r:], self.loss_ring[:r]))
