"""Run only pangu_mlp_ln_residual (single-kernel path) at a given grid; development aid for timing ablations
(PANGU_B200_GEMM_DEBUG bits with the development library tools/bin/libpangu_b200_dev.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import engine, ops, _lib
if os.environ.get('PANGU_B200_GEMM_DEBUG') or os.environ.get('DEV_LIB'):
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libpangu_b200_dev.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "lo"
Z, H, W, C = (8, 181, 360, 192) if tag == "hi" else (8, 91, 180, 384)
dev = torch.device("cuda", 0)
ws = engine.workspace(dev, Z, H, W, C)
h = torch.bfloat16
g = lambda *s: (torch.randn(*s, device=dev) * 0.02)
w1, b1, w2, b2 = g(4 * C, C).to(h), g(4 * C), g(C, 4 * C).to(h), g(C)
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
ws.x32.normal_(); ws.x16.copy_(ws.x32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    ops.mlp_ln_residual(ws.x16, w1, b1, w2, b2, gam, bet, None, ws.x32, ws.x16w[1], Z, H, W, C, 1, 1.0, False)
for _ in range(3): run()
ts = []
for _ in range(5):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
print(tag, "debug", os.environ.get('PANGU_B200_GEMM_DEBUG'), "ms %.4f" % ts[2], flush=True)
