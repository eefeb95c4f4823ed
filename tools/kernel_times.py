"""Per-kernel timing of the hot path at the full 0.25 degree shapes (CUDA events, torch's
current stream, 3 warm-ups, L2 flushed between launches).  Prints one line per entry point
with achieved TFLOP/s and algorithmic GB/s.  Development aid; bench.py is the contract."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pangu_pytorch_b200 as pb
from pangu_pytorch_b200 import engine, ops


def timeit(fn, iters=5, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    fmt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    pb.set_operand_dtype(fmt)
    fp16 = fmt == "fp16"
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for tag, (Z, H, W, C, heads) in {"hi": (8, 181, 360, 192, 6), "lo": (8, 91, 180, 384, 12)}.items():
        ws = engine.workspace(dev, Z, H, W, C)
        T, Tp, types = ws.T, ws.Tp, ws.types
        h = ops.dtype16(fp16)
        g = lambda *s: (torch.randn(*s, device=dev) * 0.02)
        w_qkv, b_qkv = g(3 * C, C).to(h), g(3 * C)
        w_o, b_o = g(C, C).to(h), g(C)
        w1, b1, w2, b2 = g(4 * C, C).to(h), g(4 * C), g(C, 4 * C).to(h), g(C)
        gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ebias = g(1, types, heads, 144, 144)
        ws.x32.normal_()
        ws.x16.copy_(ws.x32)
        ops.to_window16(ws.x32, ws.x16w[0], Z, H, W, C, 0, fp16)
        ops.to_window16(ws.x32, ws.x16w[1], Z, H, W, C, 1, fp16)

        def rec(name, ms, flops, bytes_):
            rows.append({"kernel": f"{tag}.{name}", "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                         "gbs": round(bytes_ / ms / 1e6, 1)})
            print(rows[-1], flush=True)

        ms = timeit(lambda: ops.qkv(ws.x16w[0], w_qkv, b_qkv, ws.qkv, Z, H, W, C, fp16), flush=flush)
        rec("qkv", ms, 2.0 * Tp * C * 3 * C, Tp * C * 2 + Tp * 3 * C * 2)
        for roll in (0, 1):
            ms = timeit(lambda: ops.window_attention(ws.qkv, ebias, ws.att, Z, H, W, C, heads, roll, fp16), flush=flush)
            rec(f"attention.roll{roll}", ms, 4.0 * (Tp // 144) * heads * 144 * 144 * 32,
                Tp * 3 * C * 2 + Tp * C * 2 + ebias.numel() * 4)
        ms = timeit(lambda: ops.proj_ln_residual(ws.att, w_o, b_o, gam, bet, ws.x32, ws.x16, Z, H, W, C, 1, 1.0, fp16),
                    flush=flush)
        rec("proj_ln_res", ms, 2.0 * Tp * C * C, Tp * C * 2 + T * C * (4 + 4 + 2))
        # the two MLP GEMMs separately (pangu_linear) and fused entry point
        hid = ws.hidden if ws.hidden is not None else torch.empty(T, 4 * C, dtype=ws.x16.dtype, device=dev)
        ms = timeit(lambda: ops.linear(ws.x16, w1, b1, None, hid, True, fp16), flush=flush)
        rec("mlp1_gelu", ms, 2.0 * T * C * 4 * C, T * C * 2 + T * 4 * C * 2)
        ms = timeit(lambda: ops.mlp_ln_residual(ws.x16, w1, b1, w2, b2, gam, bet, hid, ws.x32, ws.x16w[1], Z, H, W,
                                                C, 1, 1.0, fp16), flush=flush)
        rec("mlp_ln_res(both)", ms, 16.0 * T * C * C, T * C * 2 + 2 * T * 4 * C * 2 + T * C * (4 + 4 + 2))
        if True:          # single-kernel Mlp (both resolutions): taken when no hidden workspace is passed
            ms = timeit(lambda: ops.mlp_ln_residual(ws.x16, w1, b1, w2, b2, gam, bet, None, ws.x32, ws.x16w[1], Z, H, W,
                                                    C, 1, 1.0, fp16), flush=flush)
            rec("mlp_ln_res(one kernel)", ms, 16.0 * T * C * C, T * C * 2 + T * C * (4 + 4 + 2))
        ms = timeit(lambda: ops.to_window16(ws.x32, ws.x16w[0], Z, H, W, C, 0, fp16), flush=flush)
        rec("to_window16", ms, 0.0, T * C * 4 + Tp * C * 2)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/kernel_times_{fmt}.json", "w") as fh:
        json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
