"""Run only the window-attention kernel at a given grid (debug / profiling aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import engine, ops, _lib

if os.environ.get('ATTN_TRACE') or os.environ.get('PANGU_B200_ATTN_DEBUG'):     # development build with the trace hooks compiled in (tools/bin, see tools/README.md)
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libpangu_b200_dev.so")

tag = sys.argv[1] if len(sys.argv) > 1 else "hi"
W = int(sys.argv[2]) if len(sys.argv) > 2 else (360 if tag == "hi" else 180)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
Z, H, C, heads = (8, 181, 192, 6) if tag == "hi" else (8, 91, 384, 12)
dev = torch.device("cuda", 0)
ws = engine.workspace(dev, Z, H, W, C)
ws.qkv.normal_()
ebias = torch.randn(1, ws.types, heads, 144, 144, device=dev) * 0.02
trace = None
if os.environ.get('ATTN_TRACE'):
    # device memory by default (a store to pinned host memory per event slows the traced warp down by ~1.5x and with it
    # the whole CTA); ATTN_TRACE=host keeps the buffer readable even after a device trap
    trace = (torch.zeros(8 * 64 * 4, dtype=torch.int64).pin_memory() if os.environ['ATTN_TRACE'] == 'host'
             else torch.zeros(8 * 64 * 4, dtype=torch.int64, device=dev))
    os.environ['PANGU_B200_ATTN_TRACE'] = str(trace.data_ptr())
def run():
  for roll in (0, 1):
      ops.window_attention(ws.qkv, ebias, ws.att, Z, H, W, C, heads, roll, False)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(iters):
          ops.window_attention(ws.qkv, ebias, ws.att, Z, H, W, C, heads, roll, False)
      e1.record()
      torch.cuda.synchronize()
      print(tag, W, "roll", roll, "ms %.4f" % (e0.elapsed_time(e1) / iters), "mean|out| %.5f" % float(ws.att.float().abs().mean()),
            "debug", os.environ.get("PANGU_B200_ATTN_DEBUG"), "per", os.environ.get("PANGU_B200_ATTN_PER"), flush=True)

try:
    run()
except Exception as e:
    print('FAILED:', str(e)[:100])

if trace is not None:
    tr = trace.cpu().view(8, 64, 4)
    t0 = int(tr[tr > 0].min())
    names = {0: "TMA  [slot free]", 1: "MMA  [full, S issued, pfull, oempty]", 2: "TAIL [start, done]",
             3: "EXP warp [enter, S + max ready, exp done, P published]", 4: "EXP warps [publish time, by lane quadrant]", 5: "MAX warp [enter, S ready, max posted, epilogue(i-2) done]", 6: "SEG  [start, bias staged, tail-warp0 done, all done]"}
    names[7] = "MAX warps [post time, by lane quadrant]"
    for role in range(8):
        print(names[role])
        for i in range(0, 64):
            row = [int(x) - t0 if int(x) > 0 else -1 for x in tr[role, i]]
            print("   win %2d: %s" % (i, row))
