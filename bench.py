#!/usr/bin/env python
"""bench.py — frames/s of the Critic/Hourglass hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload all|hourglass|critic_train|infer]

BASELINE.json's metric is "Hourglass+critic train frames/s & mask-infer frames/s": the default run (`--workload all`)
measures all three workloads and prints ONE JSON line whose headline (`value`, `e2e`, `roofline`, `cpu_baseline`) is
the critic-guided Hourglass training step (configs[2]: frozen critic, inject, L1, batch 1024 per GPU — the largest
single-GPU configuration) and whose `workloads` object carries the same four fields for the critic training step
(configs[1], batch 256) and `-process` mask inference (configs[0], batch 256, threshold 0.1); `infer_sweep` holds the
batch 1k-64k inference sweep (configs[4]).  A "step" is one pass of the workload over one synthetic batch.

  value : whole-job frames/s with inputs resident in HBM: blocks of K back-to-back steps (CUDA-graph replays over a
          rotation of distinct resident batches larger than the 126 MB L2), one CUDA-event pair per block, the block
          repeated until the timed region is >= 200 ms; the MEDIAN block is reported (`repeats`, `timed_region_ms`)
  e2e   : the same steps through the public API from pinned HOST uint8 frames: H2D copy of every step's inputs and a D2H
          read of every step's result inside the timed region (double-buffered pipeline, >= 250 ms, independent of K)
`--impl reference` times the reference's own classes (oracle/_ref: unmodified nets.py, loop bodies of main.py) on the
host cores for the same configs; under torchrun only rank 0 works.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "hourglass_critic_train_and_mask_infer_frames_per_s"
UNIT = "frames/s"
WORKLOADS = {
    "hourglass": "critic-guided Hourglass training step, frozen critic, inject, L1 0.5, batch 1024/GPU (BASELINE configs[2])",
    "critic_train": "critic training step, batch 256/GPU, synthetic 64x64x3 uint8 frames + sparse-reward labels (BASELINE configs[1])",
    "infer": "mask inference (-process path, threshold 0.1), batch 256/GPU (BASELINE configs[0])",
}
DEFAULT_BATCH = {"hourglass": 1024, "critic_train": 256, "infer": 256}
# SURVEY.md §8d / BASELINE.md: algorithmic FLOPs per frame (2*MAC, conv+linear), chfak 1 | 5
FLOPS = {"critic_train": {1: 8460480, 5: 140729280}, "hourglass": {1: 67944832, 5: 665240704},
         "infer": {1: 20959296, 5: 186601792}}
MIN_REGION_S = 0.2
L2_BYTES = 126 * 2 ** 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    which="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1590.0, which="fallback (B200_PROFILING.md)")


def config_of(workload, B, world, chfak):
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOADS[workload], "batch_per_gpu": B, "global_batch": B * world, "chfak": chfak,
            "parallelism": f"dp{world}", "frame": "64x64x3 uint8",
            "l2": "steps rotate over distinct resident batches whose total size exceeds the 126 MB L2"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.perf_counter()

    def stop(self, windows=()):
        """Median SM clock over the samples that fall inside the timed windows [(t0, t1), ...] (all samples if none do)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if any(a <= t <= b for a, b in windows)]
        rows = inside or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "samples_in_timed_regions": len(inside)}


def make_batch(workload, B, seed):
    import cgs_b200.synth as synth
    X, Y, _ = synth.synthetic_frames(2 * B if workload == "hourglass" else B, seed=seed)
    return X, Y


# --------------------------------------------------------------------------- reference arm (CPU)
def reference_step_fn(workload, chfak, B):
    """One step of the workload on the host CPU -> (step(), kind, description).  kind "reference": the unmodified
    reference classes from oracle/_ref (nets.NewCritic / nets.UnetDecoder) driven by the loop bodies of main.py
    (critic_pipe :185-200, segmentation_training :344-463, segment :1139-1164); kind "port": the oracle restatement
    (oracle/torch_ref.py) when oracle/_ref is not present.  Synthetic data and weights of the same recipe as the GPU arm."""
    import torch.nn.functional as F
    import cgs_b200.synth as synth
    from oracle import ref_shims, torch_ref
    torch.manual_seed(0)
    p = 0.3
    csd = synth.perturbed_state(synth.critic_shapes(chfak), 0)
    msd = synth.perturbed_state(synth.masker_shapes(chfak), 1)
    X, Y = make_batch(workload, B, 0)
    Xt = torch.from_numpy(X)
    Yt = torch.from_numpy(Y[1, :B]).float()
    if ref_shims.available():
        nets, _ = ref_shims.load()
        critic = nets.NewCritic(bottleneck=32, chfak=chfak, dropout=p)          # main.py:108
        masker = nets.UnetDecoder(bottleneck=32, chfak=chfak)                   # main.py:109
        critic.load_state_dict({k: torch.from_numpy(v) for k, v in csd.items()})
        masker.load_state_dict({k: torch.from_numpy(v) for k, v in msd.items()})
        kind, desc = "reference", "unmodified reference nets.py classes (oracle/_ref), main.py loop body, torch %s CPU fp32" % torch.__version__
        if workload == "critic_train":
            critic.train()
            opt = torch.optim.Adam(critic.parameters())                         # main.py:178

            def step():
                XP = Xt.permute(0, 3, 1, 2).float() / 255.0                     # main.py:189
                pred = critic(XP).squeeze()
                loss = F.mse_loss(pred, Yt)                                     # main.py:195
                opt.zero_grad(); loss.backward(); opt.step()                    # main.py:196-198
                return loss.item()
        elif workload == "hourglass":
            critic.train(); masker.train()
            opt = torch.optim.Adam(masker.parameters())                         # main.py:334 (-frozen)

            def step():
                A = Xt[:B].permute(0, 3, 1, 2).float() / 255.0                  # main.py:360-361
                Bf = Xt[B:].permute(0, 3, 1, 2).float() / 255.0
                pred, embeds = critic(A, collect=True)                          # main.py:364
                negpred = critic(Bf)
                pred = pred.squeeze(); negpred = negpred.squeeze().detach()
                Z = masker(A, embeds)                                           # main.py:391
                replaced = A * (1 - Z) + Z * Bf                                 # main.py:395
                loss = F.mse_loss(critic(replaced).squeeze(), negpred.detach())
                injected = Bf * (1 - Z) + Z * A                                 # main.py:406
                loss = loss + F.mse_loss(critic(injected).squeeze(), pred.detach())
                loss = loss + 0.5 * F.l1_loss(Z, torch.zeros_like(Z))           # main.py:422 (staticnorm)
                opt.zero_grad(); loss.backward(); opt.step()                    # main.py:460-462
                return loss.item()
        else:
            critic.eval(); masker.eval()                                        # main.py:1128-1129

            def step():
                batch = torch.from_numpy(X / 255.0).permute(0, 3, 1, 2).float()  # main.py:1134-1136
                pred, embeds = critic(batch, collect=True)                      # main.py:1139 (autograd on, as the reference)
                mask = masker(batch, embeds)                                    # main.py:1150
                hard = mask >= 0.1                                              # main.py:1164
                return float(hard.sum())
        return step, kind, desc
    csd = {k: torch.from_numpy(v).requires_grad_(True) for k, v in csd.items()}
    msd = {k: torch.from_numpy(v).requires_grad_(True) for k, v in msd.items()}
    c = chfak
    drop = lambda: tuple(torch.nn.functional.dropout(torch.ones(s), p, True)
                         for s in ((B, 8 * c, 8, 8), (B, 16 * c, 4, 4), (B, 32 * c)))
    if workload == "critic_train":
        opt = torch.optim.Adam(csd.values())

        def step():
            XP = Xt.permute(0, 3, 1, 2).float() / 255.0
            loss, _ = torch_ref.critic_loss(csd, XP, Yt, masks=drop())
            opt.zero_grad(); loss.backward(); opt.step()
            return loss.item()
    elif workload == "hourglass":
        opt = torch.optim.Adam(msd.values())

        def step():
            A = Xt[:B].permute(0, 3, 1, 2).float() / 255.0
            Bf = Xt[B:].permute(0, 3, 1, 2).float() / 255.0
            loss, _, _ = torch_ref.hourglass_losses(csd, msd, A, Bf, Yt, live=False, inject=True, L1=0.5,
                                                    masks=[drop() for _ in range(4)])
            opt.zero_grad(); loss.backward(); opt.step()
            return loss.item()
    else:
        def step():
            batch = (torch.from_numpy(X / 255.0)).permute(0, 3, 1, 2).float()
            pred, mask, hard = torch_ref.segment_batch(csd, msd, batch, 0.1)   # autograd on, as main.py:1139
            return float(mask.detach().numpy().sum())
    return step, "port", "oracle/torch_ref.py restatement, torch %s CPU fp32" % torch.__version__


def cpu_arm(workload, chfak, B, steps, warmup, budget_s=None):
    """Times the CPU arm: `steps` steps (or as many as fit `budget_s`, at least 2) of batch B on all host cores."""
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    step, kind, desc = reference_step_fn(workload, chfak, B)
    for _ in range(warmup):
        step()
    n, t0 = 0, time.perf_counter()
    while n < steps and (budget_s is None or n < 2 or time.perf_counter() - t0 < budget_s):
        step(); n += 1
    dt = time.perf_counter() - t0
    return {"value": B * n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{n} full steps of batch {B} in {dt:.1f} s ({desc})"}, 1e3 * dt / n, n


def run_reference(args, rank, world):
    if rank != 0:
        return
    names = list(WORKLOADS) if args.workload == "all" else [args.workload]
    recs = {}
    for name in names:
        B = args.batch or DEFAULT_BATCH[name]
        # a step of this arm = one full batch of the workload; the number of timed steps is bounded so that the whole
        # run stays within a few minutes on the host (hourglass at batch 1024 is ~1-2 s per step on 16 cores)
        budget = 60.0 if name == "hourglass" else 25.0
        cb, ms, n = cpu_arm(name, args.chfak, B, args.steps, min(args.warmup, 2), budget_s=budget)
        recs[name] = {"value": cb["value"], "unit": UNIT, "ms_per_step": ms, "steps_timed": n, "config": config_of(name, B, world, args.chfak),
                      "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    head = recs[names[0]]
    line = {"metric": METRIC if args.workload == "all" else args.workload + "_frames_per_s", "value": head["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "config": head["config"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"],
            "steps_timed": head["steps_timed"]}
    if len(names) > 1:
        line["workloads"] = {k: v for k, v in recs.items() if k != names[0]}
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm (GPU)
def kernel_flops(entry, B, chfak=1):
    """Algorithmic FLOPs (2*MAC) one launch of a C-ABI entry point performs on B frames, chfak 1.
    Layer MACs per frame: features.0 884,736; .3 589,824; .6 147,456; .10 73,728; .14 + crit 9,248;
    dec[4] 1,024; dec[3] 110,592; dec[2] 110,592; dec[1] 294,912; dec[0] 1,179,648; masker.0 6,488,064; masker.2 589,824."""
    if chfak != 1:
        # the wide path (cgs_b200/wide.py): per-launch average over the launches one step makes of each entry point
        k = chfak
        l0, l123 = 884736 * k, (589824 + 147456 + 73728) * k * k
        head = (8192 + 1024) * k * k
        per = {"cgs_wide_conv3x3": 2 * l123 / 6.0,                     # features.3/.6/.10 forward + input gradients: 6 launches
               "cgs_wide_wgrad3x3": l123 / 3.0, "cgs_wide_conv0_fwd": l0, "cgs_wide_conv0_wgrad": l0, "cgs_wide_gemm": 3 * head / 6.0}
        m = per.get(entry)
        return None if m is None else 2 * m * B
    cf = 884736 + 589824 + 147456 + 73728 + 9248                       # critic forward
    cb = cf - 884736                                                   # its dgrads (no input gradient for features.0) ...
    dec = 1024 + 110592 + 110592 + 294912 + 1179648
    msk = 6488064 + 589824
    per = {
        "cgs_critic_train_fused": cf + cb + cf,                        # fprop + dgrad + wgrad = 8,460,480 / 2
        "cgs_critic_train_bf16": cf + cb + cf,                         # the same step with bf16 operands (csrc/hg_critic.cu)
        "cgs_critic_loss_xgrad": cf + cf,                              # fprop + full input gradient (incl. features.0 dgrad)
        "cgs_critic_forward_frames": cf,
        "cgs_infer_fused": cf + dec,
        "cgs_masker_fused": msk,
        "cgs_hg_forward": cf + dec + msk,                              # critic(A, collect) + decoder + masker
        "cgs_hg_score": 2 * (cf + cf),                                 # replaced + injected: fprop + input gradient each
        "cgs_hg_score_bf16": cf + 2 * (cf + cf),                       # critic(B) + replaced + injected (fprop + input gradient each)
        "cgs_hg_backward": 2 * (dec + msk) - 884736 * 0,               # wgrad + dgrad of decoder and masker
    }
    m = per.get(entry)
    return None if m is None else 2 * m * B


def time_blocks(run_block, K, sync, world, dev):
    """R blocks of K steps, one CUDA-event pair per block, R chosen so that the timed region is >= MIN_REGION_S.
    Returns (median block seconds [max over ranks], R, total seconds, wall window)."""
    import torch.distributed as dist
    sync()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run_block(); e.record()
    sync()
    est = max(s.elapsed_time(e) * 1e-3, 1e-6)
    R = int(min(max(1, -(-MIN_REGION_S // est)), 4000))
    if world > 1:
        r = torch.tensor([R], device=dev)
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
        R = int(r.item())
    evs = []
    sync()
    w0 = time.perf_counter()
    for _ in range(R):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run_block(); e.record()
        evs.append((s, e))
    sync()
    w1 = time.perf_counter()
    t = torch.tensor([a.elapsed_time(b) * 1e-3 for a, b in evs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # per block: the slowest rank
    return float(t.median().item()), R, float(t.sum().item()), (w0, w1)


def bench_workload(name, args, rank, world, dev, group, sampler_windows):
    """One workload on this rank's GPU -> record dict (rank 0's is printed)."""
    import torch.distributed as dist
    from cgs_b200 import ops
    from cgs_b200.graph_step import (GraphedCriticStep, GraphedHourglassStep, GraphedSegment, HostPipeline,
                                     PipelinedCriticTrainer, _capture, _train_state)
    from cgs_b200.train_handler import Handler, parse_args
    B, W, K = args.batch or DEFAULT_BATCH[name], max(args.warmup, 3), args.steps
    ops.set_precision(args.precision)
    hargs = parse_args(["--chfak", str(args.chfak)] + (["-frozen"] if name == "hourglass" else []))
    torch.manual_seed(0)
    H = Handler(hargs, device=dev, rank=rank, world_size=world, process_group=group)
    H.critic.to(dev); H.masker.to(dev)
    X, Y = make_batch(name, B, seed=rank)
    sync = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize
    frame_bytes = B * 12288 * (2 if name == "hourglass" else 1)
    nrot = max(2, -(-int(1.3 * L2_BYTES) // frame_bytes))                # distinct resident batches: > L2 in total
    # steps captured per graph launch: a 52 us critic step is shorter than the host's launch-to-launch jitter, and on several
    # GPUs every step waits for the slowest rank's launch; a launch of up to 20 steps keeps the host out of the step
    spg = 1 if name == "hourglass" else max(d for d in (20, 16, 10, 8, 5, 4, 2, 1) if K % d == 0)
    nrot = -(-nrot // spg) * spg
    Yb = Y[1, :B].astype(np.float32)
    launches_per_step = 0
    if name == "critic_train":
        Xrot = torch.stack([torch.from_numpy(np.roll(X, 7 * r + 1, axis=0)) for r in range(nrot)]).to(dev)
        Yrot = torch.stack([torch.from_numpy(np.roll(Yb, 7 * r + 1)) for r in range(nrot)]).to(dev)
        base = GraphedCriticStep(H, B)
        opti, launches_per_step = base.opti, base.launches
        roll0 = torch.zeros(1, dtype=torch.int32, device=dev)
        graphs = [_capture(lambda r0=r0: torch.stack([H.critic_step(Xrot[r0 + k], Yrot[r0 + k], opti, roll=roll0) for k in range(spg)]),
                           warmup=1, state=_train_state(opti, H.critic))[0] for r0 in range(0, nrot, spg)]
        trains = True
    elif name == "hourglass":
        Xrot = torch.stack([torch.from_numpy(np.roll(X, 7 * r + 1, axis=0)) for r in range(nrot)]).to(dev)
        Yrot = torch.stack([torch.from_numpy(np.roll(Yb, 7 * r + 1)) for r in range(nrot)]).to(dev)
        steps = [GraphedHourglassStep(H, B, X=Xrot[0, :B], CX=Xrot[0, B:], Y=Yrot[0])]
        opti = steps[0].opti
        steps += [GraphedHourglassStep(H, B, opti, X=Xrot[r, :B], CX=Xrot[r, B:], Y=Yrot[r], warmup=1) for r in range(1, nrot)]
        graphs, launches_per_step, trains = [s.graph for s in steps], steps[0].launches, True
    else:
        Xrot = torch.stack([torch.from_numpy(np.roll(X, 7 * r + 1, axis=0)) for r in range(nrot)]).to(dev)
        H.critic.eval(); H.masker.eval()
        segs = [GraphedSegment(H, B, 0.1, X=Xrot[r]) for r in range(nrot)]
        graphs, launches_per_step, trains, spg = [s.graph for s in segs], segs[0].launches, False, 1
    cursor = [0]

    def run_block():
        for _ in range(K // spg):
            graphs[cursor[0] % len(graphs)].replay()
            cursor[0] += 1
    Kb = (K // spg) * spg
    for _ in range(max(1, -(-W // spg))):
        graphs[cursor[0] % len(graphs)].replay(); cursor[0] += 1
    block_s, R, region_s, win = time_blocks(run_block, K, sync, world, dev)
    sampler_windows.append(win)
    if trains:
        ops.weights_changed()
        opti.check()
    value = world * B * Kb / block_s
    rec = {"value": value, "unit": UNIT, "ms_per_step": 1e3 * block_s / Kb, "repeats": R, "timed_region_ms": 1e3 * region_s,
           "steps_per_block": Kb, "config": config_of(name, B, world, args.chfak),
           "timing": (f"median of {R} blocks of {Kb} back-to-back steps (one cuda-event pair per block, {spg} step(s) per graph "
                      f"launch) over {nrot} distinct resident batches ({nrot * frame_bytes >> 20} MiB of frames > 126 MB L2)"),
           "launches_per_step": launches_per_step, "gpu_launches": launches_per_step * Kb * R,
           "achieved_tflops": FLOPS[name].get(args.chfak, 0) * value / 1e12}

    # ---- e2e: pinned host buffers -> H2D -> step -> D2H of the step's result, every step, through the public API
    if name == "critic_train":
        trainer = PipelinedCriticTrainer(H, B)
        chunk = 16                                        # steps per H2D copy / graph launch (48 MB at batch 256): fixed
        nb = 2 * chunk
        Xds = torch.from_numpy(np.concatenate([X] * nb)).pin_memory()
        Yds = torch.from_numpy(np.tile(Yb, nb)).pin_memory()
        run_e2e = lambda: trainer.train(Xds, Yds, chunk=chunk)        # nb steps
        finish = lambda: trainer.losses()
        steps_per_call, h2d, d2h = nb, B * 12288 + B * 4, 4
    elif name == "hourglass":
        slots = [GraphedHourglassStep(H, B, opti, warmup=1) for _ in range(2)]
        pipe = HostPipeline(slots, lambda s: s.out)
        Xh = [torch.from_numpy(np.roll(X, 3 * r, axis=0)).pin_memory() for r in range(4)]
        Yh = torch.from_numpy(Yb).pin_memory()

        def run_e2e():
            for r in range(4):
                pipe.step(Xh[r][:B], Xh[r][B:], Yh)
        finish = lambda: pipe.results()
        steps_per_call, h2d, d2h = 4, 2 * B * 12288 + B * 4, pipe.d2h_bytes
    else:
        slots = [GraphedSegment(H, B, 0.1) for _ in range(2)]
        pipe = HostPipeline(slots, lambda s: s.out[2])                # the thresholded masks (uint8), as -process writes them
        Xh = [torch.from_numpy(np.roll(X, 3 * r, axis=0)).pin_memory() for r in range(8)]

        def run_e2e():
            for r in range(8):
                pipe.step(Xh[r])
        finish = lambda: pipe.results()
        steps_per_call, h2d, d2h = 8, B * 12288, pipe.d2h_bytes
    run_e2e(); finish(); sync()
    t0 = time.perf_counter(); run_e2e(); finish(); est = max(time.perf_counter() - t0, 1e-5)
    calls = int(min(max(2, -(-0.25 // est)), 20000))
    if world > 1:
        c = torch.tensor([calls], device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        calls = int(c.item())
    sync()
    t0 = time.perf_counter()
    for _ in range(calls):
        run_e2e()
    res = finish()
    sync()
    e2e_s = time.perf_counter() - t0
    sampler_windows.append((t0, t0 + e2e_s))
    assert torch.isfinite(res.float()).all()
    if trains:
        ops.weights_changed()
        opti.check()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_steps = calls * steps_per_call
    rec["e2e"] = {"value": world * B * e2e_steps / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h, "steps": e2e_steps, "seconds": float(t.item()),
                  "how": "double-buffered pinned-host pipeline through cgs_b200.graph_step; every step's inputs cross PCIe and "
                         "every step's result is read back inside the timed region; step count fixed by time, not by --steps"}

    # ---- DP proof: parameters bit-identical on every rank after the timed regions
    if world > 1 and trains:
        flat = opti.flat.clone()
        ck = torch.stack([flat.double().sum(), flat.double().abs().sum(), (flat.view(torch.int32).long() % 65521).sum().double()])
        cks = [torch.zeros_like(ck) for _ in range(world)]
        dist.all_gather(cks, ck)
        rec["dp_params_equal"] = bool(all(torch.equal(cks[0], c) for c in cks))
        rec["dp_param_checksums"] = [float(c[2]) for c in cks]
        assert rec["dp_params_equal"], f"{name}: parameters differ across ranks after training: {cks}"

    # ---- per-kernel shares of one eager step (CUDA events around every C-ABI launch) -> roofline of the dominant kernel
    if rank == 0 and not args.no_extras:
        pk = peaks()
        Xd, Yd = Xrot[0], (Yrot[0] if name != "infer" else None)

        def eager():
            if name == "critic_train":
                H.critic_step(Xd, Yd, opti, roll=roll0)
            elif name == "hourglass":
                H.segmentation_step(Xd[:B], Xd[B:], Yd, opti)
            else:
                with torch.no_grad():
                    H.segment_device(Xd, 0.1)
        for _ in range(2):
            eager()
        with ops.profile_calls() as prof:
            for _ in range(5):
                # keep the GPU backlogged while the host enqueues the step: an event pair around a launch then spans the kernel's
                # execution, not the host's launch latency (which would weigh a 27-launch step's small kernels far too heavily)
                torch.cuda._sleep(int(8e6))
                eager()
        summ = prof.summary()
        tot = sum(v[1] for v in summ.values()) or 1.0
        ks = sorted(({"entry": k, "launches_per_step": n / 5.0, "ms_per_step": ms / 5.0, "share": ms / tot,
                      "flops_per_launch": kernel_flops(k, B, args.chfak)} for k, (n, ms) in summ.items()), key=lambda d: -d["ms_per_step"])
        rec["kernels"] = [{a: (round(b, 5) if isinstance(b, float) else b) for a, b in k.items()} for k in ks[:8]]
        top = ks[0]
        # live duration of the dominant kernel inside the timed region = its share of the (graph-replayed) step
        live_us = top["share"] * rec["ms_per_step"] * 1e3 / max(top["launches_per_step"], 1e-9)
        fl = top["flops_per_launch"]
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mma_tf32 = sms * 512 * 2 * 1.965e9 / 1e12         # mma.sync m16n8k8 TF32: 2 clk/instr/SM (tools/mma_rate.cu)
        tj_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
        tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
        if fl:
            tfl = fl / (live_us * 1e-6) / 1e12
            rec["roofline"] = {"bound": "tensor", "achieved": tfl, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                               "frac": tfl / pk["tf_sustained"], "traffic": tj.get(f"{top['entry']}@{B}"),
                               "kernel": top["entry"], "share_of_step": top["share"], "algorithmic_flops": fl,
                               "launch_us": live_us,
                               "peak_source": pk["which"] + ": bf16_tflops_sustained (dense tcgen05 bf16, kernel timed inside a long "
                                              "step); " + ("the kernel here is a TMA-fed tcgen05 bf16 convolution with N = 40..80 per MMA "
                                                           "(issue-bound: DESIGN.md 4e)" if args.chfak != 1 else
                                                           "no TF32 peak is measured on this pool - the kernels here are mma.sync "
                                                           "TF32/bf16, whose own issue peak is given below"),
                               "mma_sync_tf32_peak_tflops": mma_tf32, "frac_of_mma_sync_tf32_peak": tfl / mma_tf32,
                               "mma_sync_bf16_peak_tflops": 2 * mma_tf32, "frac_of_mma_sync_bf16_peak": tfl / (2 * mma_tf32),
                               "step_frac_of_peak": rec["achieved_tflops"] / world / pk["tf_sustained"]}
        else:
            by = B * 12288 * 4
            rec["roofline"] = {"bound": "hbm", "achieved": by / (live_us * 1e-6) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                               "frac": by / (live_us * 1e-6) / 1e9 / pk["hbm"], "traffic": None, "kernel": top["entry"],
                               "share_of_step": top["share"], "launch_us": live_us, "peak_source": pk["which"]}
        if world == 1:
            cb, _, _ = cpu_arm(name, args.chfak, B, 10 ** 6, 1, budget_s=12.0)
            rec["cpu_baseline"] = cb
    return rec


def infer_sweep(args, dev):
    """BASELINE configs[4]: mask inference at batch 1k-64k resident frames (one launch pair per batch), frames/s."""
    from cgs_b200 import ops
    from cgs_b200.train_handler import Handler, parse_args
    import cgs_b200.synth as synth
    torch.manual_seed(0)
    H = Handler(parse_args(["--chfak", str(args.chfak)]), device=dev)
    H.critic.to(dev).eval(); H.masker.to(dev).eval()
    X, _, _ = synth.synthetic_frames(4096, seed=0)
    out = []
    for n in (1024, 4096, 16384, 65536):
        Xd = torch.from_numpy(X).to(dev).repeat(n // 4096 if n >= 4096 else 1, 1, 1, 1)[:n].contiguous()
        with torch.no_grad():
            for _ in range(2):
                H.segment_device(Xd, 0.1)
            torch.cuda.synchronize()
            reps = max(2, 65536 // n)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                H.segment_device(Xd, 0.1)
            e.record()
            torch.cuda.synchronize()
        out.append({"batch": n, "frames_per_s": n * reps / (s.elapsed_time(e) * 1e-3)})
        del Xd
    return out


def run_ours(args, rank, world):
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    names = list(WORKLOADS) if args.workload == "all" else [args.workload]
    sampler = ClockSampler(local)
    windows = []
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    recs = {n: bench_workload(n, args, rank, world, dev, group, windows) for n in names}
    wide_rec = None
    if args.workload == "all" and args.chfak == 1 and world == 1 and not args.no_extras and args.precision == "tf32":
        # the paper's width (chfak 5, reference docs/index.html:151) through the TMA / tcgen05 kernels of the wide path
        import copy
        a5 = copy.copy(args)
        a5.chfak = 5
        wide_rec = bench_workload("critic_train", a5, rank, world, dev, group, windows)
    clocks = sampler.stop(windows) if rank == 0 else None
    if rank == 0:
        head = recs[names[0]]
        line = {"metric": METRIC if args.workload == "all" else names[0] + "_frames_per_s", "value": head["value"], "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": ("bf16" if names[0] in ("hourglass", "infer") or os.environ.get("CGS_CRITIC_BF16", "1") != "0" else "tf32")
                if args.precision == "tf32" else "f32",
                "data": "synthetic"}
        line.update({k: v for k, v in head.items() if k not in ("value", "unit", "ms_per_step")})
        line["precision"] = ("whole-frame kernels: bf16 mma.sync operands (fp32 accumulate) in every 3x3 convolution of the Hourglass "
                             "step, of -process inference and of the critic training step (CGS_CRITIC_BF16=0: TF32 mma.sync "
                             "there); 4x4 / 1x1 / Linear layers, losses, Adam fp32" if args.precision == "tf32" else "all fp32 (FFMA)")
        line["clocks"] = clocks
        if len(names) > 1:
            line["workloads"] = {k: v for k, v in recs.items() if k != names[0]}
            line["gpu_launches"] = sum(r["gpu_launches"] for r in recs.values())
        if wide_rec is not None:
            line["workloads"]["critic_train_chfak5"] = wide_rec
        if world == 1 and not args.no_extras and ("infer" in names):
            line["infer_sweep"] = infer_sweep(args, dev)
        print(json.dumps(line))
    if world > 1:
        # Tearing the communicator down while captured graphs still reference it hangs in ncclCommAbort on this
        # stack (tools/dp_check.py); all ranks are done, so flush and leave without destroy_process_group().
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config of each workload)")
    ap.add_argument("--chfak", type=int, default=1)
    ap.add_argument("--no-extras", action="store_true", help="skip per-kernel roofline, CPU baseline and sweep legs")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"],
                    help="tf32: tensor-core kernels (TF32 / bf16 operands, fp32 accumulate); fp32: exact FFMA kernels")
    args = ap.parse_args()
    if args.workload == "all" and args.batch:
        raise SystemExit("bench.py: --batch needs a single --workload")
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
