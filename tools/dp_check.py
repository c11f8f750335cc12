"""Staged data-parallel check (run under torchrun on >= 2 GPUs): prints after every stage so a hang is attributable."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get("DP_CHECK_TIMEOUT", "90")), exit=True)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    sys.stdout.write(f"[rank {rank} t={time.time()-T0:5.1f}] " + " ".join(str(x) for x in a) + "\n")      # one write: lines of two ranks do not interleave
    sys.stdout.flush()
T0 = time.time()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
say("pg up"); dist.barrier(); torch.cuda.synchronize(); say("barrier ok")
from cgs_b200 import ops
from cgs_b200.train_handler import Handler, parse_args, FlatAdam
from cgs_b200.graph_step import GraphedCriticStep
import cgs_b200.synth as synth
ops.set_precision(os.environ.get("DP_PRECISION", "tf32"))
torch.manual_seed(0)
H = Handler(parse_args(["--dropout", "0"]), device=dev, rank=rank, world_size=world, process_group=dist.group.WORLD)
H.critic.to(dev)
B = 64
X, Y, _ = synth.synthetic_frames(B * world, seed=0)
Xs = torch.from_numpy(X[rank * B:(rank + 1) * B]).to(dev); Ys = torch.from_numpy(Y[1, rank * B:(rank + 1) * B]).float().to(dev)
opt = FlatAdam(H.critic.parameters(), process_group=dist.group.WORLD, world_size=world)
for i in range(3):
    l = H.critic_step(Xs, Ys, opt)
torch.cuda.synchronize(); say("eager DP steps ok, loss", float(l), "| p2p all-reduce:", opt._p2p is not None, "ok:", opt.p2p_ok())
if opt._p2p is not None:      # same three steps with the NCCL all-reduce instead of the peer-memory kernel
    os.environ["CGS_P2P"] = "0"
    torch.manual_seed(0)
    Hn = Handler(parse_args(["--dropout", "0"]), device=dev, rank=rank, world_size=world, process_group=dist.group.WORLD)
    Hn.critic.to(dev)
    on = FlatAdam(Hn.critic.parameters(), process_group=dist.group.WORLD, world_size=world)
    os.environ["CGS_P2P"] = "1"
    for i in range(3):
        Hn.critic_step(Xs, Ys, on)
    torch.cuda.synchronize()
    say(f"p2p vs NCCL: max |dparam| = {(on.flat - opt.flat).abs().max().item():.3e}")
flat = opt.flat.clone(); ref = flat.clone(); dist.broadcast(ref, 0); torch.cuda.synchronize()
say("params equal across ranks:", bool(torch.equal(flat, ref)))
if world > 1 and rank == 0:
    # single-process reference on the global batch
    torch.manual_seed(0)
    H1 = Handler(parse_args(["--dropout", "0"]), device=dev); H1.critic.to(dev)
    o1 = FlatAdam(H1.critic.parameters())
    Xg = torch.from_numpy(X).to(dev); Yg = torch.from_numpy(Y[1, :B * world]).float().to(dev)
    for i in range(3):
        H1.critic_step(Xg, Yg, o1)
    torch.cuda.synchronize()
    err = (o1.flat - flat).abs().max().item()
    say(f"DP vs single-process global batch: max |dparam| = {err:.3e} (scale {o1.flat.abs().max().item():.3e})")
dist.barrier(); say("capturing graph")
step = GraphedCriticStep(H, B, opt)
torch.cuda.synchronize(); say("captured, launches", step.launches)
step.load(Xs.cpu().pin_memory(), Ys.cpu().pin_memory())
for i in range(5):
    out = step.replay()
torch.cuda.synchronize(); say("replays ok, loss", float(out))
ref = opt.flat.clone(); dist.broadcast(ref, 0); torch.cuda.synchronize()
say("params equal across ranks after graph replays:", bool(torch.equal(opt.flat, ref)))
# ---- the frozen-critic Hourglass step, data-parallel (cgs_p2p_stage + cgs_p2p_allreduce_adam behind the whole-frame kernels)
dist.barrier()
torch.manual_seed(1)
Hh = Handler(parse_args(["-frozen", "--dropout", "0"]), device=dev, rank=rank, world_size=world, process_group=dist.group.WORLD)
Hh.critic.to(dev).train(); Hh.masker.to(dev).train()
for q in Hh.critic.parameters():
    q.requires_grad_(False)
oh = FlatAdam(list(Hh.masker.parameters()), process_group=dist.group.WORLD, world_size=world)
Bh = 32
Xh, _, _ = synth.synthetic_frames(2 * Bh * world, seed=5)
A_all, C_all = Xh[:Bh * world], Xh[Bh * world:]
Ah = torch.from_numpy(A_all[rank * Bh:(rank + 1) * Bh]).to(dev); Ch = torch.from_numpy(C_all[rank * Bh:(rank + 1) * Bh]).to(dev)
for i in range(3):
    t = Hh.segmentation_step(Ah, Ch, None, oh)
torch.cuda.synchronize()
fh = oh.flat.clone(); rh = fh.clone(); dist.broadcast(rh, 0); torch.cuda.synchronize()
say("hourglass params equal across ranks:", bool(torch.equal(fh, rh)))
if world > 1 and rank == 0:
    torch.manual_seed(1)
    H1 = Handler(parse_args(["-frozen", "--dropout", "0"]), device=dev)
    H1.critic.to(dev).train(); H1.masker.to(dev).train()
    for q in H1.critic.parameters():
        q.requires_grad_(False)
    o1 = FlatAdam(list(H1.masker.parameters()))
    Ag, Cg = torch.from_numpy(A_all).to(dev), torch.from_numpy(C_all).to(dev)
    for i in range(3):
        H1.segmentation_step(Ag, Cg, None, o1)
    torch.cuda.synchronize()
    err = (o1.flat - fh).abs().max().item()
    say(f"hourglass DP vs single-process global batch: max |dparam| = {err:.3e} (scale {o1.flat.abs().max().item():.3e})")
dist.barrier(); torch.cuda.synchronize(); say("done"); sys.stdout.flush(); os._exit(0)
