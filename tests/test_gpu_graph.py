"""GPU: the CUDA-graph captured steps (cgs_b200/graph_step.py) follow the SAME trajectory as eager calls of the loop
bodies from the same initial state: constructing a graphed step must not advance parameters, Adam moments, the step count
or the dropout stream (the capture warm-up used to apply real optimizer steps on zero-filled buffers)."""
import numpy as np
import pytest
import torch

import cgs_b200.synth as synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _handler(argv, seed_c=41, seed_m=42):
    from cgs_b200.train_handler import Handler, parse_args
    torch.manual_seed(0)
    H = Handler(parse_args(argv), device=DEV)
    H.critic.load_state_dict({k: torch.from_numpy(v) for k, v in synth.perturbed_state(synth.critic_shapes(1), seed_c, 1.5).items()})
    H.masker.load_state_dict({k: torch.from_numpy(v) for k, v in synth.perturbed_state(synth.masker_shapes(1), seed_m, 1.5).items()})
    H.critic.to(DEV); H.masker.to(DEV)
    return H


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_graphed_critic_steps_match_eager(precision):
    from cgs_b200 import ops
    from cgs_b200.graph_step import GraphedCriticStep, PipelinedCriticTrainer
    B, N = 64, 6
    X, Y, _ = synth.synthetic_frames(B * N, seed=7)
    Xh, Yh = torch.from_numpy(X), torch.from_numpy(Y[1]).float()
    ops.set_precision(precision)
    try:
        # eager trajectory
        H = _handler(["--dropout", "0.3"])
        H.critic.train()
        opt = H._opt(H.critic.parameters())
        le = [H.critic_step(Xh[i * B:(i + 1) * B].to(DEV), Yh[i * B:(i + 1) * B].to(DEV), opt, roll=3).item() for i in range(N)]
        pe = opt.flat.clone()
        # graphed: one replay per step
        H2 = _handler(["--dropout", "0.3"])
        g = GraphedCriticStep(H2, B)
        assert int(g.opti.step_count[0].item()) == 0, "capture warm-up leaked optimizer steps"
        g.roll.fill_(3)
        lg = [g(Xh[i * B:(i + 1) * B], Yh[i * B:(i + 1) * B]).item() for i in range(N)]
        pg = g.opti.flat.clone()
        assert int(g.opti.step_count[0].item()) == N
        # tf32 = the whole-step kernel (fixed-order partial vectors: bit-reproducible); fp32 = per-layer kernels whose weight
        # gradients arrive by float REDs (summation order varies run to run)
        ptol = 1e-7 if precision == "tf32" else 2e-6
        np.testing.assert_allclose(lg, le, rtol=1e-6 if precision == "tf32" else 1e-5, atol=1e-9)
        assert (pg - pe).abs().max().item() <= ptol * max(pe.abs().max().item(), 1.0)
        # pipelined trainer: chunked path (chunk 2) + ragged tail through step()
        H3 = _handler(["--dropout", "0.3"])
        tr = PipelinedCriticTrainer(H3, B)
        assert int(tr.opti.step_count[0].item()) == 0
        rolls = [3] * N
        tr.train(Xh[:5 * B].pin_memory(), Yh[:5 * B].pin_memory(), rolls=rolls[:5], chunk=2)
        tr.step(Xh[5 * B:].pin_memory(), Yh[5 * B:].pin_memory(), roll=3)
        lp = tr.losses().numpy()
        np.testing.assert_allclose(lp, le, rtol=1e-6 if precision == "tf32" else 1e-5, atol=1e-9)
        assert (tr.opti.flat - pe).abs().max().item() <= ptol * max(pe.abs().max().item(), 1.0)
    finally:
        ops.set_precision("fp32")


def test_pipelined_trainer_ring_wrap_keeps_every_loss():
    from cgs_b200 import ops
    from cgs_b200.graph_step import PipelinedCriticTrainer
    B, N = 16, 12
    X, Y, _ = synth.synthetic_frames(B * N, seed=8)
    Xh, Yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y[1]).float().pin_memory()
    ops.set_precision("tf32")
    try:
        H = _handler(["--dropout", "0"])
        tr = PipelinedCriticTrainer(H, B, ring=8)
        tr.train(Xh, Yh, chunk=3)                      # 4 chunks of 3: the third wraps the ring of 8
        H2 = _handler(["--dropout", "0"])
        tr2 = PipelinedCriticTrainer(H2, B, ring=64)
        tr2.train(Xh, Yh, chunk=3)
        np.testing.assert_allclose(tr.losses().numpy(), tr2.losses().numpy()[-8:], rtol=0, atol=0)
    finally:
        ops.set_precision("fp32")


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_graphed_hourglass_steps_match_eager(precision):
    from cgs_b200 import ops
    from cgs_b200.graph_step import GraphedHourglassStep
    B, N = 24, 3
    X, Y, _ = synth.synthetic_frames(2 * B * N, seed=9)
    ops.set_precision(precision)
    try:
        H = _handler(["-frozen", "--dropout", "0.3"])
        H.critic.train(); H.masker.train()
        for q in H.critic.parameters():
            q.requires_grad_(False)
        opt = H._opt(H.masker.parameters())
        te = []
        for i in range(N):
            xs = X[2 * B * i:2 * B * (i + 1)]
            t = H.segmentation_step(xs[:B], xs[B:], torch.from_numpy(Y[1, 2 * B * i:2 * B * i + B]), opt, roll=-2)
            te.append([t[k].item() for k in sorted(t)])
        pe = opt.flat.clone()
        H2 = _handler(["-frozen", "--dropout", "0.3"])
        g = GraphedHourglassStep(H2, B)
        assert int(g.opti.step_count[0].item()) == 0
        g.roll.fill_(-2)
        tg = []
        for i in range(N):
            xs = torch.from_numpy(X[2 * B * i:2 * B * (i + 1)])
            tg.append(g(xs[:B], xs[B:], torch.from_numpy(Y[1, 2 * B * i:2 * B * i + B]).float()).tolist())
        np.testing.assert_allclose(tg, te, rtol=2e-5, atol=1e-8)     # gradient REDs of the per-layer path are order-dependent
        assert (g.opti.flat - pe).abs().max().item() <= 2e-5 * max(pe.abs().max().item(), 1.0)
    finally:
        ops.set_precision("fp32")
