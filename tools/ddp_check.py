"""Multi-GPU check of the native training path under stock DistributedDataParallel (one process per GPU, NCCL):
each rank trains on its own frame; after backward every rank must hold the SAME gradients (DDP's all-reduce ran on
the gradients our autograd Function returned) and they must equal the mean of the per-rank gradients computed
without DDP.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
from bnerv_b200 import HNeRV_Boost, tiny_args  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
a = tiny_args("HNeRV_Boost")
torch.manual_seed(3)
model = HNeRV_Boost(a).to(dev).train()
ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
fh, fw = [int(v) for v in a.fc_hw.split("_")]
g = torch.Generator().manual_seed(100 + rank)
emb = torch.rand(1, 16, fh, fw, generator=g).to(dev)
t = torch.tensor([(rank + 1) / 8], dtype=torch.float64, device=dev)
target = torch.rand(1, 3, fh * 20, fw * 20, generator=g).to(dev)


def loss_of(m):
    img = m(None, emb, norm_idx=t)[0]
    return ((img - target) ** 2).mean()


# local gradients without DDP
model.zero_grad(set_to_none=True)
loss_of(model).backward()
local_g = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
mean_g = {}
for n, gr in local_g.items():
    s = gr.clone()
    dist.all_reduce(s)
    mean_g[n] = s / world
# through DDP
model.zero_grad(set_to_none=True)
loss_of(ddp).backward()
worst = 0.0
for n, p in model.named_parameters():
    if n in mean_g:
        worst = max(worst, ((p.grad - mean_g[n]).abs().max() / mean_g[n].abs().max().clamp_min(1e-30)).item())
w = torch.tensor([worst], device=dev)
dist.all_reduce(w, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"ddp_check: world {world}, {len(mean_g)} parameters, max rel diff DDP grads vs mean of per-rank grads {w.item():.2e}")
    assert w.item() < 1e-5
dist.destroy_process_group()
