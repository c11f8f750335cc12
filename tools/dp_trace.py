"""Phase timing of the data-parallel whole-step kernel's tail (run under torchrun on >= 2 GPUs): partial write + grid barrier,
local slice sum, publish + wait for the peers' flags, peer loads + Adam."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from cgs_b200 import ops, _lib
from cgs_b200.train_handler import Handler, parse_args, FlatAdam
ops.set_precision("tf32")
B = 256
torch.manual_seed(0)
H = Handler(parse_args([]), device=dev, rank=rank, world_size=world, process_group=dist.group.WORLD)
H.critic.to(dev).train()
opt = FlatAdam(H.critic.parameters(), process_group=dist.group.WORLD, world_size=world)
X = torch.randint(0, 255, (B, 64, 64, 3), dtype=torch.uint8, device=dev); Y = torch.rand(B, device=dev)
L = _lib.lib()
for _ in range(5):
    H.critic_step(X, Y, opt)
torch.cuda.synchronize(); dist.barrier()
tr = torch.zeros(4 * 24 + 4 * 160, dtype=torch.int64, device=dev)
sync = torch.zeros(1, device=dev)
for rep in range(3):
    tr.zero_(); dist.all_reduce(sync)
    L.cgs_critic_fused_set_trace(tr.data_ptr()); H.critic_step(X, Y, opt); torch.cuda.synchronize(); L.cgs_critic_fused_set_trace(None)
    T = tr.cpu()[:96].view(4, 24); C = tr.cpu()[96:].view(160, 4); C = C[C[:, 0] != 0]
    E = [T[f] for f in range(4) if T[f][16] != 0][0]
    print(f"[rank {rank}] kernel {int(C[:,2].max() - C[:,0].min())} ns | tail clk: write+barrier {int(E[18]-E[17])}  "
          f"slice sum + push + poll + Adam {int(E[19]-E[18])}  rest {int(E[16]-E[19])}", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
