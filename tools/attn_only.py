"""Run only the window-attention kernel at a given grid (debug / profiling aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import engine, ops

tag = sys.argv[1] if len(sys.argv) > 1 else "hi"
W = int(sys.argv[2]) if len(sys.argv) > 2 else (360 if tag == "hi" else 180)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
Z, H, C, heads = (8, 181, 192, 6) if tag == "hi" else (8, 91, 384, 12)
dev = torch.device("cuda", 0)
ws = engine.workspace(dev, Z, H, W, C)
ws.qkv.normal_()
ebias = torch.randn(1, ws.types, heads, 144, 144, device=dev) * 0.02
for roll in (0, 1):
    for _ in range(iters):
        ops.window_attention(ws.qkv, ebias, ws.att, Z, H, W, C, heads, roll, False)
    torch.cuda.synchronize()
    print(tag, W, "roll", roll, "ok", float(ws.att.float().abs().mean()), flush=True)
