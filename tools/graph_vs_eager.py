"""Full 0.25 degree forward: eager launches vs CUDA-graph replay (device time per step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pangu_pytorch_b200 as pb
from pangu_pytorch_b200.rollout import GraphedStep

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pb.PanguModel(device=dev).to(dev).eval()
g = torch.Generator(device=dev).manual_seed(1)
up = torch.randn(1, 5, 13, 721, 1440, device=dev, generator=g)
sf = torch.randn(1, 4, 721, 1440, device=dev, generator=g)
maps = torch.randn(1, 3, 724, 1440, device=dev, generator=g)
ch = torch.randn(1, 1, 1, 13, 721, 1440, device=dev, generator=g)
stats = [torch.zeros(4, device=dev), torch.ones(4, device=dev), torch.zeros(13, 1, 1, 5, device=dev), torch.ones(13, 1, 1, 5, device=dev)]


def timed(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad():
    eager = timed(lambda: model(up, sf, stats, maps, ch))
    step = GraphedStep(model, up, sf, stats, maps, ch)
    graphed = timed(lambda: step.graph.replay())
print(f"eager {eager:.3f} ms/step, graph replay (forward + denorm) {graphed:.3f} ms/step")
