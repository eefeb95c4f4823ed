"""Run one GEMM entry point (qkv | proj) at a given grid; development aid: timing and, with GEMM_TRACE=1 and the development
library (tools/bin/libpangu_b200_dev.so), the clock64 timeline of one CTA (PANGU_B200_GEMM_DEBUG >> 8 selects it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import engine, ops, _lib
if os.environ.get('PANGU_B200_GEMM_DEBUG') or os.environ.get('DEV_LIB') or os.environ.get('GEMM_TRACE'):
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libpangu_b200_dev.so")
what = sys.argv[1] if len(sys.argv) > 1 else "qkv"
tag = sys.argv[2] if len(sys.argv) > 2 else "lo"
Z, H, W, C = (8, 181, 360, 192) if tag == "hi" else (8, 91, 180, 384)
dev = torch.device("cuda", 0)
ws = engine.workspace(dev, Z, H, W, C)
h = torch.bfloat16
g = lambda *s: (torch.randn(*s, device=dev) * 0.02)
w_qkv, b_qkv, w_o, b_o = g(3 * C, C).to(h), g(3 * C), g(C, C).to(h), g(C)
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
ws.x32.normal_(); ws.x16.copy_(ws.x32); ws.att.normal_()
ops.to_window16(ws.x32, ws.x16w[0], Z, H, W, C, 0, False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
trace = None
if os.environ.get('GEMM_TRACE'):
    trace = torch.zeros(8 * 64 * 4, dtype=torch.int64, device=dev)
    os.environ['PANGU_B200_GEMM_TRACE'] = str(trace.data_ptr())
def run():
    if what == "qkv":
        ops.qkv(ws.x16w[0], w_qkv, b_qkv, ws.qkv, Z, H, W, C, False)
    else:
        ops.proj_ln_residual(ws.att, w_o, b_o, gam, bet, ws.x32, ws.x16, Z, H, W, C, 1, 1.0, False)
for _ in range(3): run()
ts = []
for _ in range(5):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
print(what, tag, "debug", os.environ.get('PANGU_B200_GEMM_DEBUG'), "ms %.4f" % ts[2], flush=True)
if trace is not None:
    tr = trace.cpu().view(8, 64, 4)
    t0 = int(tr[0, 0, 0])
    rel = lambda v: int(v) - t0 if int(v) else -1
    names = ["producer, k-block: slot free, loads issued", "issuer, k-block: before wait, full seen, MMAs + commit issued",
             "issuer, tile: before accumulator wait, accumulator free", "epilogue warp 2, tile: before wait, accumulator full, drained (before arrive)",
             "epilogue warp 6, tile: same", "", "", ""]
    for role in range(5):
        print("role", role, names[role])
        for i in range(64):
            if int(tr[role, i].abs().sum()) == 0:
                continue
            print("  %2d " % i + " ".join("%8d" % rel(v) for v in tr[role, i]))
