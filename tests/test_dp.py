"""CPU, world_size 2 over gloo: the data-parallel host logic — batch sharding, 1/world loss scaling, the
single flat-bucket all-reduce in FlatAdam, identical parameters on every rank and equality with a
single-process step on the global batch.  The CUDA Adam kernel is replaced by a test-only torch stand-in
(the product itself has no CPU path); gradients come from the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_adam(p, g_buf, m, v, step_state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, clear_grad=False):
    step_state[0] += 1
    t = int(step_state[0].item())
    g = g_buf * grad_scale
    m.lerp_(g, 1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    denom = (v.sqrt() / (1 - betas[1] ** t) ** 0.5).add_(eps)
    p.addcdiv_(m, denom, value=-lr / (1 - betas[0] ** t))
    if clear_grad:
        g_buf.zero_()


def _grads(critic, X, Y, scale):
    from oracle import torch_ref
    sd = dict(critic.named_parameters())
    loss, _ = torch_ref.critic_loss(sd, torch_ref.to_input(X), Y)
    gs = torch.autograd.grad(loss * scale, list(sd.values()))
    return loss.detach(), gs


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from cgs_b200 import ops
    from cgs_b200.train_handler import FlatAdam, Handler, parse_args
    import cgs_b200.synth as synth
    ops.adam_step = _cpu_adam                      # test-only stand-in for cgs_adam_step
    torch.manual_seed(0)
    H = Handler(parse_args(["--dropout", "0"]), device="cpu", rank=rank, world_size=world, process_group=dist.group.WORLD)
    X, Y, _ = synth.synthetic_frames(16, seed=1)
    Yt = torch.from_numpy(Y[1]).float()
    opt = FlatAdam(H.critic.parameters(), process_group=dist.group.WORLD, world_size=world)
    keys = list(H.critic.state_dict().keys())
    for step in range(3):
        sl = H._shard(len(X))
        opt.zero_grad()
        _, gs = _grads(H.critic, X[sl], Yt[sl], 1.0 / world)
        for p, g in zip(H.critic.parameters(), gs):
            p.grad.add_(g)                          # what the wgrad kernels do into the bucket views
        opt.step()
    flat = opt.flat.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        # single-process run on the global batch
        torch.manual_seed(0)
        H1 = Handler(parse_args(["--dropout", "0"]), device="cpu")
        o1 = FlatAdam(H1.critic.parameters())
        for step in range(3):
            o1.zero_grad()
            _, gs = _grads(H1.critic, X, Yt, 1.0)
            for p, g in zip(H1.critic.parameters(), gs):
                p.grad.add_(g)
            o1.step()
        out.put(dict(same=bool(all(torch.equal(gathered[0], t) for t in gathered)),
                     err=float((flat - o1.flat).abs().max()), scale=float(o1.flat.abs().max()),
                     views=bool(all(H.critic.state_dict()[k].data_ptr() >= opt.flat.data_ptr() for k in keys)),
                     shards=[(H._shard(16).start, H._shard(16).stop)]))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["same"], "ranks diverged after the all-reduced update"
    assert res["views"], "parameters are no longer views of the flat bucket"
    assert res["err"] <= 2e-6 * max(res["scale"], 1.0), res


def test_shards_partition_the_batch():
    """Balanced, never-empty shards (ceil(n/world) left ranks empty on ragged DataLoader tails) whose
    B_local/B_global weights make the rank-summed gradient the global-batch mean."""
    import sys
    sys.path.insert(0, ROOT)
    from cgs_b200.train_handler import Handler, parse_args
    for world in (1, 2, 4, 8):
        hs = [Handler(parse_args([]), device="cpu", rank=r, world_size=world) for r in range(world)]
        for n in (64, 128, 8192, 10, 65, 129, 9, 6, 5, 3, 2, 1):
            sls = [h._shard(n) for h in hs]
            ws = [h._shard_weight(n) for h in hs]
            assert all(sl.stop > sl.start for sl in sls), (world, n, sls)
            assert abs(sum(ws) - 1.0) < 1e-12, (world, n, ws)
            if n >= world:
                covered = [i for sl in sls for i in range(sl.start, sl.stop)]
                assert covered == list(range(n))
                sizes = [sl.stop - sl.start for sl in sls]
                assert max(sizes) - min(sizes) <= 1
            else:
                assert all((sl.start, sl.stop) == (0, n) for sl in sls)


def _ragged_worker(rank, world, port, out):
    """Ragged global batch (n = 13 over 2 ranks, then n = 1): weighted shards must reproduce the single-process update."""
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from cgs_b200 import ops
    from cgs_b200.train_handler import FlatAdam, Handler, parse_args
    import cgs_b200.synth as synth
    ops.adam_step = _cpu_adam
    torch.manual_seed(100 + rank)                  # deliberately different seeds: FlatAdam must broadcast rank 0's parameters
    from cgs_b200.nets import NewCritic
    critic = NewCritic(dropout=0.0)
    opt = FlatAdam(critic.parameters(), process_group=dist.group.WORLD, world_size=world)
    H = Handler.__new__(Handler)
    H.rank, H.world = rank, world
    X, Y, _ = synth.synthetic_frames(13, seed=4)
    Yt = torch.from_numpy(Y[1]).float()
    init = opt.flat.clone()
    for n in (13, 1):
        sl, w = H._shard(n), H._shard_weight(n)
        opt.zero_grad()
        _, gs = _grads(critic, X[:n][sl], Yt[:n][sl], w)
        for p, g in zip(critic.parameters(), gs):
            p.grad.add_(g)
        opt.step()
    flats = [torch.zeros_like(opt.flat) for _ in range(world)]
    inits = [torch.zeros_like(init) for _ in range(world)]
    dist.all_gather(flats, opt.flat.clone())
    dist.all_gather(inits, init)
    if rank == 0:
        c1 = NewCritic(dropout=0.0)
        o1 = FlatAdam(c1.parameters())
        o1.flat.copy_(init)
        for n in (13, 1):
            o1.zero_grad()
            _, gs = _grads(c1, X[:n], Yt[:n], 1.0)
            for p, g in zip(c1.parameters(), gs):
                p.grad.add_(g)
            o1.step()
        out.put(dict(same_init=bool(all(torch.equal(inits[0], t) for t in inits)),
                     same=bool(all(torch.equal(flats[0], t) for t in flats)),
                     err=float((opt.flat - o1.flat).abs().max()), scale=float(o1.flat.abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_ragged_batch_and_broadcast_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_ragged_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["same_init"], "FlatAdam did not broadcast rank 0's initial parameters"
    assert res["same"], "ranks diverged on a ragged batch"
    assert res["err"] <= 2e-6 * max(res["scale"], 1.0), res
