"""2-rank data-parallel equivalence on GPUs (run under torchrun): gradients of a 2-rank run on 2 x B images must equal
the single-process gradients on the same 2B images (no BatchNorm), SURVEY.md §4(iv)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-unet_b200"))
import torch
import torch.distributed as dist

import b200unet
from b200unet.ddp import DataParallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
model = b200unet.UNet(1, 2, 3, 6, False, False, "upconv").cuda()
B = 2
g = torch.Generator().manual_seed(5)
X = torch.randn(world * B, 1, 92, 92, generator=g).cuda()
Y = torch.randint(0, 2, (world * B, 52, 52), generator=g).cuda()
# single-process reference on the whole batch (every rank computes it)
loss = model.loss(X, Y)
loss.backward()
ref = {n: p.grad.clone() for n, p in model.named_parameters()}
model.zero_grad(set_to_none=True)
ddp = DataParallel(model, bucket_bytes=1 << 20)
loss = ddp.loss(X[rank * B:(rank + 1) * B], Y[rank * B:(rank + 1) * B])
loss.backward()
torch.cuda.synchronize()
worst = 0.0
for n, p in model.named_parameters():
    e = float((p.grad - ref[n]).norm() / ref[n].norm().clamp_min(1e-30))
    worst = max(worst, e)
print(f"rank {rank}: buckets {len(ddp.bucketer.buckets)} worst rel-L2 vs single-process grads {worst:.3e}", flush=True)
assert worst < 2e-2, worst
dist.destroy_process_group()
