"""Run only pangu_mlp_ln_residual (single-kernel path) at a given grid; development aid for timing ablations
(PANGU_B200_GEMM_DEBUG bits with the development library tools/bin/libpangu_b200_dev.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import engine, ops, _lib
if os.environ.get('PANGU_B200_GEMM_DEBUG') or os.environ.get('DEV_LIB') or os.environ.get('MLP_TRACE'):
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
trace = None
if os.environ.get('MLP_TRACE'):          # clock64 timeline of one CTA (PANGU_B200_GEMM_DEBUG >> 8 selects it), device memory
    trace = torch.zeros(8 * 64 * 4, dtype=torch.int64, device=dev)
    os.environ['PANGU_B200_MLP_TRACE'] = str(trace.data_ptr())
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

if trace is not None:
    tr = trace.cpu().view(8, 64, 4)
    t0 = int(tr[0, 0, 0])
    rel = lambda v: int(v) - t0 if int(v) else -1
    names = ["W1 load: slot free, issued", "W2 load: h0 free, h0 issued, h1 free, h1 issued", "G1: Hacc free, W1 full, committed",
             "G2: H full, W2 h0 full, W2 h1 full, committed", "GELU wg0: Hacc full, math done, H buffer free, arrived",
             "GELU wg1: Hacc full, math done, H buffer free, arrived", "epilogue warp 6: [tile*8] Y full | block i: start, resid full, stored, drained",
             "epilogue warp 6, block i: accumulator in registers, math + staging written, fence done, TMA store issued"]
    for role in range(8):
        print("role", role, names[role])
        for g in range(56):
            if int(tr[role, g].abs().sum()) == 0:
                continue
            print("  %2d " % g + " ".join("%8d" % rel(v) for v in tr[role, g]))
