"""Full 0.25 degree training step (forward on the tape + weighted-L1 loss + backward) on one B200:
wall time per phase and per-entry-point device time (CUDA events).  Development aid."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import ops, training

LAT, LON = 721, 1440


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = pb.PanguModel(device=dev).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(1)
    up = torch.randn(1, 5, 13, LAT, LON, device=dev, generator=g)
    sf = torch.randn(1, 4, LAT, LON, device=dev, generator=g)
    maps = torch.randn(1, 3, 724, LON, device=dev, generator=g)
    ch = torch.randn(1, 1, 1, 13, LAT, LON, device=dev, generator=g)
    tu = torch.randn(1, 5, 13, LAT, LON, device=dev, generator=g)
    ts = torch.randn(1, 4, LAT, LON, device=dev, generator=g)
    stats = [torch.zeros(4, device=dev), torch.ones(4, device=dev), torch.zeros(13, 1, 1, 5, device=dev),
             torch.ones(13, 1, 1, 5, device=dev)]

    def step():
        for p in model.parameters():
            p.grad = None
        return training.train_step(model, up, sf, stats, maps, ch, tu, ts)

    for _ in range(2):
        loss = step()
    torch.cuda.synchronize()
    print("loss", float(loss), "mem GB", torch.cuda.max_memory_allocated() / 2**30)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 3
    e[0].record()
    for _ in range(n):
        step()
    e[1].record()
    torch.cuda.synchronize()
    ms = e[0].elapsed_time(e[1]) / n
    prof = ops.EventProfile()
    ops.set_profile(prof)
    step()
    ops.set_profile(None)
    kern = prof.summary()
    tot = sum(v[0] for v in kern.values())
    rows = {k: {"ms": round(v[0], 3), "calls": v[1]} for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
    out = {"train_step_ms": round(ms, 2), "profiled_kernel_ms": round(tot, 2), "entry_points": rows,
           "finite_grads": all(torch.isfinite(p.grad).all().item() for p in model.parameters()),
           "algorithmic_tflop": 25.26, "tflops": round(25.26 / (ms * 1e-3), 1)}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/train_times.json", "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
