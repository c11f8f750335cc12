#!/usr/bin/env python
"""bench.py — frames/s of the Critic/Hourglass hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload critic_train|hourglass|infer]

Default workload (BASELINE.json configs[1], the configuration the metric is quoted on): one critic
training step (uint8->float, NewCritic fwd with dropout, MSE, backward, Adam) on synthetic 64x64x3
frames with sparse-reward labels, batch 256 per GPU (weak scaling).  A "step" is one such pass.
  value : whole-job frames/s with inputs resident in HBM (CUDA-graph replay, CUDA-event timed,
          L2 flushed between timed iterations)
  e2e   : the same step through the public API with pinned HOST uint8 frames + labels: H2D copies
          and a D2H read of the loss inside the timed region
`--impl reference` times the reference's CPU path (oracle port, all host threads) on the same config.
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

METRIC = "critic_train_frames_per_s"
WORKLOADS = {
    "critic_train": "critic training step, batch 256/GPU, synthetic 64x64x3 uint8 frames + sparse-reward labels (BASELINE configs[1])",
    "hourglass": "critic-guided Hourglass step, frozen critic, inject, L1, batch 1024/GPU (BASELINE configs[2])",
    "infer": "mask inference (-process path, threshold 0.1), batch 256/GPU (BASELINE configs[0])",
}
# SURVEY.md §8d / BASELINE.md: algorithmic FLOPs per frame (2*MAC, conv+linear), chfak 1 | 5
FLOPS = {"critic_train": {1: 8460480, 5: 140729280}, "hourglass": {1: 67944832, 5: 665240704},
         "infer": {1: 20959296, 5: 186601792}}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"


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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_batch(workload, B, seed):
    import cgs_b200.synth as synth
    X, Y, _ = synth.synthetic_frames(2 * B if workload == "hourglass" else B, seed=seed)
    return X, Y


# --------------------------------------------------------------------------- reference arm (CPU)
def oracle_step_fn(workload, chfak, B):
    """The reference's CPU path restated by the oracle (reference classes are not on the GPU box)."""
    from oracle import torch_ref
    import cgs_b200.synth as synth
    torch.manual_seed(0)
    p = 0.3
    csd = {k: torch.from_numpy(v).requires_grad_(True) for k, v in synth.perturbed_state(synth.critic_shapes(chfak), 0).items()}
    msd = {k: torch.from_numpy(v).requires_grad_(True) for k, v in synth.perturbed_state(synth.masker_shapes(chfak), 1).items()}
    X, Y = make_batch(workload, B, 0)
    Xt = torch.from_numpy(X)
    Yt = torch.from_numpy(Y[1, :B]).float()
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
    return step


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    B = args.batch
    step = oracle_step_fn(args.workload, args.chfak, B)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    line = {"metric": METRIC if args.workload == "critic_train" else args.workload + "_frames_per_s", "value": v,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOADS[args.workload], "batch_per_step": B, "chfak": args.chfak},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.steps} full steps of batch {B} (oracle/torch_ref.py, torch {torch.__version__} CPU fp32)"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm (GPU)
def time_kernel(fn, flush, iters=20):
    """Average device time of one launch: CUDA events on the launch stream, L2 flushed in between."""
    for _ in range(3):
        fn()
    evs = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return float(np.mean([s.elapsed_time(e) for s, e in evs])) * 1e-3


def kernel_rooflines(B, chfak, flush, hbm_gbs):
    """Dominant-kernel candidates of the critic step, timed in isolation."""
    from cgs_b200 import ops
    from cgs_b200._lib import SRC_PLAIN, SRC_POOLBWD, EPI_RELU_POOL
    dev = "cuda"
    c = chfak
    out = []
    for name, H, Cin, Cout in (("features.0", 64, 3, 8 * c), ("features.3", 32, 8 * c, 8 * c)):
        x = torch.rand(B, H, H, Cin, device=dev)
        w = torch.rand(Cout, Cin, 3, 3, device=dev) - 0.5
        b = torch.zeros(Cout, device=dev)
        e = torch.empty(B, H // 2, H // 2, Cout, device=dev)
        idx = torch.empty(B, H // 2, H // 2, Cout, device=dev, dtype=torch.uint8)
        de = torch.rand_like(e)
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        t_f = time_kernel(lambda: ops.conv3x3(ops._src(SRC_PLAIN, Cin, x), w, b, B, H, H, Cout, EPI_RELU_POOL, e, idx_out=idx), flush)
        t_w = time_kernel(lambda: ops.wgrad3x3(ops._src(SRC_PLAIN, Cin, x), ops._src(SRC_POOLBWD, Cout, de, e, idx), B, H, H, dw, db), flush)
        by_f = x.numel() * 4 + e.numel() * 5 + w.numel() * 4
        by_w = x.numel() * 4 + e.numel() * 9 + w.numel() * 4
        fl = 2 * B * H * H * Cin * Cout * 9
        out.append({"kernel": f"conv {name}", "desc": "fprop+bias+ReLU+maxpool", "bytes": by_f, "flops": fl, "sec": t_f})
        out.append({"kernel": f"wgrad {name}", "desc": "weight+bias gradient", "bytes": by_w, "flops": fl, "sec": t_w})
    for k in out:
        k["gbs"] = k["bytes"] / k["sec"] / 1e9
        k["tflops"] = k["flops"] / k["sec"] / 1e12
        k["frac_hbm"] = k["gbs"] / hbm_gbs
    return out


def fused_step_roofline(B, flush, hbm_gbs, tf_peak):
    """The whole-step critic kernel (csrc/critic_fused.cu) timed alone: one graph node replayed between CUDA events, L2
    flushed in between (gradient leaves as per-CTA partial vectors; the Adam tail of the full step is not in this number).
    Algorithmic work per frame: 8,460,480 FLOP; compulsory HBM bytes 12,288 (uint8 frame) + 4 (label)."""
    from cgs_b200 import ops
    from cgs_b200.nets import NewCritic
    from cgs_b200.train_handler import FlatAdam
    torch.manual_seed(0)
    c = NewCritic(dropout=0.3).cuda().train()
    opt = FlatAdam(c.parameters())
    X = torch.randint(0, 255, (B, 64, 64, 3), dtype=torch.uint8, device="cuda")
    Y = torch.rand(B, device="cuda")
    run = lambda: ops.critic_train_fused(c, X, Y, 3, rng=c._dropout_rng(X.device))      # masks drawn in-kernel, as in the step
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        run()
    sec = time_kernel(g.replay, flush)
    flops, byts = 8460480 * B, (12288 + 4) * B
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mma_peak = sms * 512 * 2 * 1.965e9 / 1e12      # mma.sync m16n8k8 TF32: 2.0 clk per instruction per SM (tools/mma_rate.cu)
    return {"kernel": "critic_fused_train_kernel", "desc": "frame -> forward -> loss -> backward, all activations in smem",
            "sec": sec, "flops": flops, "bytes": byts, "tflops": flops / sec / 1e12, "gbs": byts / sec / 1e9,
            "frac_hbm": byts / sec / 1e9 / hbm_gbs, "frac_tensor_bf16_peak": flops / sec / 1e12 / tf_peak,
            "mma_sync_tf32_peak_tflops": mma_peak, "frac_mma_sync_tf32_peak": flops / sec / 1e12 / mma_peak}


def masker_roofline(B, flush, hbm_gbs, tf_peak):
    """Dominant kernel of the inference workload: cgs_masker_fused (masker.0 + LeakyReLU + masker.2 + sigmoid + threshold),
    timed alone.  Per frame: 2 * (6,488,064 + 589,824) FLOP; HBM bytes 12,288 (frame) + 32,768
    (o0) + 16,384 (mask) + 4,096 (hard mask)."""
    from cgs_b200 import ops
    from cgs_b200.nets import UnetDecoder
    torch.manual_seed(0)
    m = UnetDecoder().cuda().eval()
    X = torch.randint(0, 255, (B, 64, 64, 3), dtype=torch.uint8, device="cuda")
    o0 = torch.rand(B, 32, 32, 8, device="cuda")
    run = lambda: ops.masker_fused(m, X, o0, 0.1)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        run()
    sec = time_kernel(g.replay, flush)
    flops, byts = 2 * (6488064 + 589824) * B, (12288 + 32768 + 16384 + 4096) * B
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mma_peak = sms * 512 * 2 * 1.965e9 / 1e12
    return {"kernel": "masker_fused_kernel", "desc": "cat(X, ups(o0)) -> masker.0 -> LeakyReLU -> masker.2 -> sigmoid -> threshold",
            "sec": sec, "flops": flops, "bytes": byts, "tflops": flops / sec / 1e12, "gbs": byts / sec / 1e9,
            "frac_hbm": byts / sec / 1e9 / hbm_gbs, "frac_tensor_bf16_peak": flops / sec / 1e12 / tf_peak,
            "mma_sync_tf32_peak_tflops": mma_peak, "frac_mma_sync_tf32_peak": flops / sec / 1e12 / mma_peak}


def run_ours(args, rank, world):
    import torch.distributed as dist
    from cgs_b200 import ops
    from cgs_b200.graph_step import GraphedCriticStep, GraphedHourglassStep, GraphedSegment
    from cgs_b200.train_handler import Handler, parse_args
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    B, W, K = args.batch, args.warmup, args.steps
    ops.set_precision(args.precision)
    hargs = parse_args(["--chfak", str(args.chfak)] + (["-frozen"] if args.workload == "hourglass" else []))
    torch.manual_seed(0)
    H = Handler(hargs, device=dev, rank=rank, world_size=world, process_group=group)
    H.critic.to(dev); H.masker.to(dev)
    X, Y = make_batch(args.workload, B, seed=rank)
    Xh = torch.from_numpy(X).pin_memory()
    Yh = torch.from_numpy(Y[1, :B]).float().pin_memory()
    if args.workload == "critic_train":
        step = GraphedCriticStep(H, B)
        host = (Xh, Yh)
    elif args.workload == "hourglass":
        step = GraphedHourglassStep(H, B)
        host = (Xh[:B], Xh[B:], Yh)
    else:
        step = GraphedSegment(H, B, 0.1)
        host = (Xh,)
    h2d = sum(t.numel() * t.element_size() for t in host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    sync = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize

    # ---- value: inputs resident in HBM, graph replay, per-step CUDA events, L2 flushed between steps
    step.load(*host)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)          # let nvidia-smi start streaming before the (short) timed region
    for _ in range(max(W, 3)):
        step.replay()
    sync()
    evs = []
    t_wall0 = time.perf_counter()
    align = torch.zeros(1, device=dev)
    for _ in range(K):
        flush.zero_()
        if world > 1:
            # ranks drift apart by host jitter while they flush; a step can only finish when the slowest peer's gradient
            # arrives, so start the timed region of every rank at a common device-side point (a tiny all-reduce)
            dist.all_reduce(align)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); step.replay(); e.record()
        evs.append((s, e))
    sync()
    t_wall = time.perf_counter() - t_wall0
    dev_s = sum(s.elapsed_time(e) for s, e in evs) * 1e-3
    flushed_ms = 1e3 * dev_s / K
    timing = "cuda events per step, L2 flushed (256 MiB memset) between timed steps"
    if args.workload == "critic_train":
        # ---- headline timing: K steps BACK TO BACK over a rotation of distinct resident batches whose total size exceeds
        # the 126 MB L2 (every step reads frames that are not cached), one event pair around the K steps.  Per-step timing
        # with a flush in between (kept as `ms_per_step_flushed`) idles the GPU before every step, so it adds the graph
        # launch latency to each step and, on several GPUs, the skew the ranks pick up while flushing.
        from cgs_b200.graph_step import _capture
        spg = max(d for d in (8, 4, 2, 1) if K % d == 0)           # steps captured per graph (one launch runs spg steps)
        nrot = -(-max(2, -(-160 * 2 ** 20 // (B * 12288))) // spg) * spg
        Xrot = torch.stack([torch.from_numpy(np.roll(X, 7 * r + 1, axis=0)) for r in range(nrot)]).to(dev)
        Yrot = torch.stack([torch.from_numpy(np.roll(Y[1, :B], 7 * r + 1)).float() for r in range(nrot)]).to(dev)
        roll0 = torch.zeros(1, dtype=torch.int32, device=dev)

        def chunk_fn(r0):
            return lambda: torch.stack([H.critic_step(Xrot[r0 + k], Yrot[r0 + k], step.opti, roll=roll0) for k in range(spg)])
        rot = [_capture(chunk_fn(r0), warmup=1)[0] for r0 in range(0, nrot, spg)]
        for g in rot[:max(1, -(-max(W, 3) // spg))]:
            g.replay()
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(K // spg):
            rot[i % len(rot)].replay()
        e.record()
        sync()
        dev_s = s.elapsed_time(e) * 1e-3
        timing = (f"one cuda-event pair around K back-to-back steps ({spg} steps per graph launch) over {nrot} distinct resident "
                  f"batches ({nrot * B * 12288 >> 20} MiB of frames > 126 MB L2); per-step timing with an L2 flush before every "
                  f"step in ms_per_step_flushed")
    # ---- e2e: pinned host buffers -> H2D -> step -> D2H of the step's result, every step, through the public API
    if args.workload == "critic_train":
        from cgs_b200.graph_step import PipelinedCriticTrainer
        trainer = PipelinedCriticTrainer(H, B)            # chunked double-buffered H2D, async loss read-back
        # steps per chunk (= per H2D copy and per graph launch).  Measured (tools/e2e_timeline.py): a chunk costs ~0.1-0.25 ms
        # of fixed latency (cross-stream event + graph launch), so large chunks win in steady state (16 -> PCIe-bound at the
        # 43-54 GB/s the host memory feeds); short runs take smaller ones so the un-overlapped first copy stays ~1/5
        chunk = max(1, min(16, K // 5))
        nb = 2 * chunk                                     # pinned host dataset of nb batches, walked K steps in total
        Xds = torch.from_numpy(np.concatenate([X] * nb)).pin_memory()
        Yds = torch.from_numpy(np.tile(Y[1, :B], nb)).float().pin_memory()
        trainer.train(Xds, Yds, chunk=chunk)
        sync()
        n0 = trainer.i
        t0 = time.perf_counter()
        done = 0
        while done < K:
            m = min(nb, K - done)
            done += trainer.train(Xds[:m * B], Yds[:m * B], chunk=chunk)
        losses = trainer.losses()                          # synchronises: all K losses are on the host
        sync()
        e2e_s = time.perf_counter() - t0
        assert trainer.i - n0 == K and torch.isfinite(losses).all() and float(losses.max()) < 10.0 and float(losses.min()) >= 0.0
        d2h = 4
    else:
        for _ in range(3):
            out = step(*host)
        sync()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(K):
            out = step(*host)
            res = out if torch.is_tensor(out) else out[0]
            val = res.reshape(-1)[:1].cpu() if args.workload != "infer" else out[2].cpu()   # loss scalar / hard masks
            d2h = val.numel() * val.element_size()
        sync()
        e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = t.tolist()
    if rank == 0:
        hbm, tf, which = peaks()
        value = world * B * K / dev_s
        line = {"metric": METRIC if args.workload == "critic_train" else args.workload + "_frames_per_s",
                "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
                "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload], "batch_per_gpu": B, "global_batch": B * world,
                           "precision": (("whole step in one kernel: TF32 mma.sync convolutions (fprop, dgrad, wgrad), fp32 "
                                          "accumulate; head, loss, Adam fp32" if (args.workload == "critic_train" and args.chfak == 1)
                                          else "two whole-frame kernels (encoder+decoder, masker): TF32 mma.sync convolutions, fp32 "
                                          "accumulate" if (args.workload == "infer" and args.chfak == 1)
                                          else "conv fprop/dgrad: tcgen05 kind::tf32, fp32 accumulate in TMEM; wgrad: TF32 mma.sync; "
                                          "head, losses, Adam: fp32") if args.precision == "tf32" else "all fp32 (FFMA)"),
                           "chfak": args.chfak, "parallelism": f"dp{world}", "timing": timing, "graph": True},
                "ms_per_step_flushed": flushed_ms,
                "e2e": {"value": world * B * K / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h},
                "gpu_launches": step.launches * K, "launches_per_step": step.launches,   # kernels, not graph launches
                "wall_ms_per_step_incl_flush": 1e3 * t_wall / K, "clocks": clocks,
                "achieved_tflops": FLOPS[args.workload].get(args.chfak, 0) * value / 1e12}
        if world == 1 and not args.no_extras:
            tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
            tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
            fused = (args.workload == "critic_train" and args.precision == "tf32" and args.chfak == 1)
            if args.workload == "infer" and args.precision == "tf32" and args.chfak == 1:
                pj = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if which == "measured" else {}
                tf_burst = pj.get("bf16_tflops", 1590.0)
                k = masker_roofline(B, flush, hbm, tf_burst)
                line["roofline"] = {"bound": "tensor", "achieved": k["tflops"], "peak": tf_burst, "unit": "TFLOP/s",
                                    "frac": k["frac_tensor_bf16_peak"], "traffic": tj.get(k["kernel"]) if B == 256 else None,
                                    "kernel": k["kernel"], "algorithmic_flops": k["flops"], "algorithmic_bytes": k["bytes"],
                                    "peak_source": which + " (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)",
                                    "launch_us": k["sec"] * 1e6, "hbm_gbs_achieved": k["gbs"], "frac_hbm": k["frac_hbm"],
                                    "mma_sync_tf32_peak_tflops": k["mma_sync_tf32_peak_tflops"],
                                    "frac_of_mma_sync_tf32_peak": k["frac_mma_sync_tf32_peak"]}
                line["kernels"] = [{kk: (round(v, 4) if isinstance(v, float) else v) for kk, v in k.items()}]
            elif fused:
                # dominant kernel = the whole-step kernel (79 % of the step, profiles/): a dense-contraction kernel whose
                # operands never leave shared memory -> tensor roofline; HBM traffic is the uint8 frames only
                pj = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if which == "measured" else {}
                tf_burst = pj.get("bf16_tflops", 1590.0)
                k = fused_step_roofline(B, flush, hbm, tf_burst)
                # the step IS this one kernel, so its launch duration is measured live over the timed region itself
                # (events around the K back-to-back steps); k = the same kernel alone, cold, without its Adam tail
                sec = dev_s / K
                tfl = k["flops"] / sec / 1e12
                pj2 = pj.get("bf16_tflops_sustained", tf_burst)
                line["roofline"] = {"bound": "tensor", "achieved": tfl, "peak": pj2, "unit": "TFLOP/s",
                                    "frac": tfl / pj2, "traffic": tj.get(k["kernel"]) if B == 256 else None,
                                    "kernel": "critic_fused_kernel<0>", "algorithmic_flops": k["flops"], "algorithmic_bytes": k["bytes"],
                                    "peak_source": which + " (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside the K-step region)",
                                    "launch_us": sec * 1e6, "hbm_gbs_achieved": k["bytes"] / sec / 1e9,
                                    "frac_hbm": k["bytes"] / sec / 1e9 / hbm,
                                    "isolated_cold_launch_us": k["sec"] * 1e6,
                                    "mma_sync_tf32_peak_tflops": k["mma_sync_tf32_peak_tflops"],
                                    "frac_of_mma_sync_tf32_peak": tfl / k["mma_sync_tf32_peak_tflops"],
                                    "note": "TF32 mma.sync m16n8k8 (N = 8 output channels rules out tcgen05 tiles); its own "
                                            "measured peak is 512 MAC/clk/SM = 0.18 of the bf16 tcgen05 peak"}
                line["kernels"] = [{kk: (round(v, 4) if isinstance(v, float) else v) for kk, v in k.items()}]
            else:
                ks = kernel_rooflines(B, args.chfak, flush, hbm)
                top = max(ks, key=lambda k: k["sec"])
                traffic = tj.get(top["kernel"]) if (args.chfak == 1 and B == 256) else None
                line["roofline"] = {"bound": "hbm", "achieved": top["gbs"], "peak": hbm, "unit": "GB/s",
                                    "frac": top["gbs"] / hbm, "traffic": traffic, "kernel": top["kernel"],
                                    "algorithmic_bytes": top["bytes"],
                                    "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
                                    "launch_us": top["sec"] * 1e6, "achieved_tflops_fp32": top["tflops"]}
                line["kernels"] = [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in kk.items()} for kk in ks]
            # CPU baseline: oracle port on the host cores, bounded sample
            cores = os.cpu_count()
            torch.set_num_threads(cores)
            ostep = oracle_step_fn(args.workload, args.chfak, B)
            ostep(); ostep()
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < 10.0 and n < 200:
                ostep(); n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": B * n / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{n} full steps of batch {B} in {dt:.1f}s (oracle/torch_ref.py on torch CPU fp32)"}
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="critic_train", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--chfak", type=int, default=1)
    ap.add_argument("--no-extras", action="store_true", help="skip per-kernel roofline and CPU baseline legs")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"],
                    help="tf32: tcgen05 TF32 conv fprop/dgrad (fp32 accumulate) where covered; fp32: exact FFMA kernels")
    args = ap.parse_args()
    if args.batch == 0:
        args.batch = 1024 if args.workload == "hourglass" else 256
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        if args.workload == "hourglass":
            args.batch = min(args.batch, 64)      # bounded sample: ~90 ms/step on CPU at the reference's own batch
        return run_reference(args, rank)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
