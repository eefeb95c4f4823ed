"""Data-parallel gradient check (run under torchrun, 2+ GPUs): every rank takes a training step on its own
sample with the overlapped GradReducer; the reduced gradients must equal the mean of the per-rank gradients,
which rank 0 recomputes locally, one sample after the other, without any collective."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import training
from pangu_pytorch_b200.dist import GradReducer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
LAT, LON = 721, 192
torch.manual_seed(0)
model = pb.PanguModel(device=dev).to(dev).train()
for m in model.modules():
    if hasattr(m, "drop_path"):
        m.drop_path.drop_prob = 0.0


def sample(r):
    g = torch.Generator(device=dev).manual_seed(100 + r)
    f = lambda *s: torch.randn(*s, device=dev, generator=g)
    return f(1, 5, 13, LAT, LON), f(1, 4, LAT, LON), f(1, 5, 13, LAT, LON), f(1, 4, LAT, LON)


g0 = torch.Generator(device=dev).manual_seed(1)
maps = torch.randn(1, 3, 724, LON, device=dev, generator=g0)
ch = torch.randn(1, 1, 1, 13, LAT, LON, device=dev, generator=g0)
stats = [torch.zeros(4, device=dev), torch.ones(4, device=dev), torch.zeros(13, 1, 1, 5, device=dev), torch.ones(13, 1, 1, 5, device=dev)]


def grads_of(r, reducer):
    model.grad_reducer = reducer
    for p in model.parameters():
        p.grad = None
    up, sf, tu, ts = sample(r)
    training.train_step(model, up, sf, stats, maps, ch, tu, ts)
    torch.cuda.synchronize()
    return [p.grad.clone() for p in model.parameters()]


reduced = grads_of(rank, GradReducer())
dist.barrier()
if rank == 0:
    local_grads = [grads_of(r, None) for r in range(world)]
    worst = 0.0
    for i, g in enumerate(reduced):
        want = sum(lg[i].double() for lg in local_grads) / world
        worst = max(worst, float((g.double() - want).norm() / want.norm().clamp_min(1e-30)))
    print(f"DDP_CHECK world={world} worst rel-L2 (reduced vs local mean) = {worst:.3e}", "OK" if worst < 1e-3 else "FAIL")
dist.barrier()
dist.destroy_process_group()
