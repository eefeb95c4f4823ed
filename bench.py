#!/usr/bin/env python
"""Benchmark of the Pangu-Weather 24 h forecast step at 0.25 degrees (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--operands bf16|fp16]

A "step" is one full PanguModel forward (embed -> 16 earth-specific blocks -> recovery) of one
synthetic 0.25 degree sample (upper 5x13x721x1440, surface 4x721x1440), batch 1, random-init
weights of the reference architecture.  N > 1 (launched by torch.distributed.run) shards
independent ensemble members over the GPUs of one box: no data-path collective, weak scaling.

One JSON line is printed by rank 0 (see the keys below).  ``value`` is device-resident
throughput; ``e2e`` is the same metric through the public ``PanguModel.forward`` with the
step's input fields in pinned host memory (H2D inside the timed region) and the forecast
fields copied back (D2H); copies are double-buffered on side streams.
``--impl reference`` times the CPU oracle port of the reference forward (the Python reference
itself cannot travel to the GPU box) on all host cores, each step a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "24h forecast steps/s @0.25deg"
UNIT = "steps/s"
FLOPS_PER_STEP = 8.4212e12          # dense-contraction FLOPs of one forward
FLOPS_BLOCKS = 8.1325e12            # attention + MLP of the 16 blocks
LAT, LON = 721, 1440
STRIP = 96                          # CPU sample: full-depth forward on a 96-column strip (1/15 of the grid)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels behind each entry point, from the committed
# `ncu --set full` captures (profiles/r01c_*.md; an entry point = the sum of its kernels)
NCU_TRAFFIC_BYTES = {
    "pangu_mlp_ln_residual[lo]": (102.0e6 + 352.3e6) + (628.9e6 + 246.0e6),     # CfgMLP1 + CfgLNRes384
    "pangu_mlp_ln_residual[hi]": None,
    "pangu_window_attention[lo]": 405.9e6 + 92.6e6,
    "pangu_qkv[lo]": 107.1e6 + 270.3e6,
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d["bf16_tflops_sustained"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference forward on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_sample_setup():
    import torch
    from oracle import pangu_oracle as O
    try:                                  # the GPU arm may have pinned this process next to its GPU: the CPU leg gets every core
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.reference_like_weights(seed=0)
    inputs = O.synthetic_inputs(seed=1, lat=LAT, lon=STRIP)
    return O, p, inputs


def cpu_sample_time(O, p, inputs) -> float:
    t0 = time.perf_counter()
    O.forward(p, *inputs)
    return time.perf_counter() - t0


def cpu_baseline(reps: int = 2) -> dict:
    import torch
    O, p, inputs = cpu_sample_setup()
    cpu_sample_time(O, p, inputs)                      # warm-up
    ts = sorted(cpu_sample_time(O, p, inputs) for _ in range(reps))
    t = ts[len(ts) // 2]
    scale = LON // STRIP
    return {"value": 1.0 / (t * scale), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"full-depth fp32 forward (oracle port of models/pangu_model.py:50-87) on a {STRIP}-column "
                      f"longitude strip = 1/{scale} of the 0.25deg grid, {t:.2f}s per sample, scaled x{scale}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    O, p, inputs = cpu_sample_setup()
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_sample_time(O, p, inputs)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_sample_time(O, p, inputs)
    dt = (time.perf_counter() - t0) / steps
    scale = LON // STRIP
    value = 1.0 / (dt * scale)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "PanguModel 24h forward, 0.25deg (721x1440x13 levels), batch 1, random-init weights",
                       "note": "CPU oracle port of the reference forward; each step is a bounded sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{STRIP}-column longitude strip (1/{scale} of the grid), full depth, "
                                       f"{dt:.2f}s per sample, scaled x{scale}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln for (t, ln) in self.lines if t0 <= t <= t1] or [ln for (_, ln) in self.lines]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(index: int) -> str:
    """Pin this process to the CPU cores next to its GPU (NVML's ideal affinity) before any pinned host buffer is
    allocated, so that the end-to-end copies of the 8 ranks do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return f"cpu affinity = NVML ideal set of GPU {index} ({len(os.sched_getaffinity(0))} cores)"
    except Exception as e:          # not fatal: the copies just may be slower
        return f"unbound ({type(e).__name__})"


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import pangu_pytorch_b200 as pb
    from pangu_pytorch_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if os.environ.get("PANGU_BENCH_BIND_NUMA") else "unbound (default; PANGU_BENCH_BIND_NUMA=1 binds to the GPU's NVML-ideal cores: measured no effect at 8 ranks)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pb.set_operand_dtype(args.operands)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- model: random-init weights of the reference architecture (same on every rank)
    torch.manual_seed(0)
    model = pb.PanguModel(device=dev).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    maps = torch.randn(1, 3, 724, LON, device=dev, generator=g)
    const_h = torch.randn(1, 1, 1, 13, LAT, LON, device=dev, generator=g)
    stats = [torch.zeros(4, device=dev), torch.ones(4, device=dev),
             torch.zeros(13, 1, 1, 5, device=dev), torch.ones(13, 1, 1, 5, device=dev)]
    # two perturbed ensemble members per rank, resident in HBM (seeds 100+k)
    members = []
    for k in range(2):
        gk = torch.Generator(device=dev).manual_seed(100 + rank * 2 + k)
        members.append((torch.randn(1, 5, 13, LAT, LON, device=dev, generator=gk),
                        torch.randn(1, 4, LAT, LON, device=dev, generator=gk)))

    def step(i):
        up, sf = members[i % 2]
        return model(up, sf, stats, maps, const_h)

    with torch.no_grad():
        for i in range(max(3, args.warmup)):
            step(i)
        barrier()
        # ---- timed region 1: device-resident inputs
        sampler = ClockSampler(local) if rank == 0 else None
        prof = ops.EventProfile()
        ops.set_profile(prof)
        l0 = ops.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ops.set_profile(None)
        launches = ops.launches() - l0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t0, t1) if sampler else None
        kern = prof.summary()

        # ---- timed region 2: end to end through the public API, host buffers, pipelined copies
        host_in = [(torch.empty(1, 5, 13, LAT, LON).pin_memory(), torch.empty(1, 4, LAT, LON).pin_memory())
                   for _ in range(2)]
        for (hu, hs), (du, ds) in zip(host_in, members):
            hu.copy_(du); hs.copy_(ds)
        host_out = [(torch.empty(1, 5, 13, LAT, LON).pin_memory(), torch.empty(1, 4, LAT, LON).pin_memory())
                    for _ in range(2)]
        dev_in = [(torch.empty_like(members[0][0]), torch.empty_like(members[0][1])) for _ in range(2)]
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        main = torch.cuda.current_stream()

        def e2e_run(nsteps):
            in_ready = [None, None]
            in_free = [None, None]
            out_done = [None, None]

            def issue_h2d(j):
                b = j % 2
                with torch.cuda.stream(s_in):
                    if in_free[b] is not None:
                        s_in.wait_event(in_free[b])
                    dev_in[b][0].copy_(host_in[b][0], non_blocking=True)
                    dev_in[b][1].copy_(host_in[b][1], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s_in); in_ready[b] = ev

            issue_h2d(0)
            for j in range(nsteps):
                b = j % 2
                if j + 1 < nsteps:
                    issue_h2d(j + 1)
                main.wait_event(in_ready[b])
                ou, os_ = model(dev_in[b][0], dev_in[b][1], stats, maps, const_h)
                ev = torch.cuda.Event(); ev.record(main); in_free[b] = ev
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev)
                    host_out[b][0].copy_(ou, non_blocking=True)
                    host_out[b][1].copy_(os_, non_blocking=True)
                    ou.record_stream(s_out); os_.record_stream(s_out)
                    d = torch.cuda.Event(); d.record(s_out); out_done[b] = d
            s_out.synchronize()

        e2e_run(2)
        barrier()
        w0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        e2e_s = time.perf_counter() - w0

    # ---- reduce over ranks (max time)
    times = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(times[0]), float(times[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    ms_per_step = ms_max / args.steps
    value = world * args.steps / (ms_max / 1e3)
    e2e_value = world * args.steps / (e2e_ms_max / 1e3)
    h2d = (5 * 13 + 4) * LAT * LON * 4
    # dominant entry point (by device time inside the timed region) and its roofline
    T_hi, T_lo = 8 * 181 * 360, 8 * 91 * 180
    Tp_hi, Tp_lo = 30 * 124 * 144, 15 * 64 * 144
    algo_flops = {
        "pangu_mlp_ln_residual[hi]": 16.0 * T_hi * 192 * 192, "pangu_mlp_ln_residual[lo]": 16.0 * T_lo * 384 * 384,
        "pangu_qkv[hi]": 6.0 * Tp_hi * 192 * 192, "pangu_qkv[lo]": 6.0 * Tp_lo * 384 * 384,
        "pangu_proj_ln_residual[hi]": 2.0 * Tp_hi * 192 * 192, "pangu_proj_ln_residual[lo]": 2.0 * Tp_lo * 384 * 384,
        "pangu_window_attention[hi]": 4.0 * 3720 * 6 * 144 * 144 * 32, "pangu_window_attention[lo]": 4.0 * 960 * 12 * 144 * 144 * 32,
    }
    shares = {k: round(v[0] / ms_max, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
    dom = max((k for k in kern if k in algo_flops), key=lambda k: kern[k][0])
    dom_ms = kern[dom][0] / kern[dom][1]
    achieved = algo_flops[dom] / (dom_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": dom, "achieved": round(achieved, 1), "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": round(achieved / pk["tflops_sustained"], 4),
                "traffic": NCU_TRAFFIC_BYTES.get(dom), "traffic_source": "ncu --set full, profiles/r01c_*.md (bytes per launch)",
                "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "avg_launch_ms": round(dom_ms, 4), "algorithmic_flops_per_launch": algo_flops[dom],
                "step_tflops": round(FLOPS_PER_STEP / (ms_per_step * 1e-3) / 1e12, 1),
                "step_frac_of_peak": round(FLOPS_PER_STEP / (ms_per_step * 1e-3) / 1e12 / pk["tflops_sustained"], 4),
                "attn_mlp_frac_of_peak": round(FLOPS_BLOCKS / (sum(v[0] for k, v in kern.items() if k in algo_flops) / args.steps * 1e-3)
                                               / 1e12 / pk["tflops_sustained"], 4),
                "kernel_time_shares": shares}
    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.operands, "data": "synthetic",
            "config": {"workload": "PanguModel 24h forward, 0.25deg (upper 5x13x721x1440, surface 4x721x1440), batch 1, "
                                   "random-init weights; one ensemble member per step per GPU",
                       "parallelism": f"ensemble members sharded over {world} GPU(s), no collective",
                       "accumulate": "fp32", "residual_stream": "fp32",
                       "l2": "per-step working set (>4 GB of activations) is far larger than the 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d,
                    "ms_per_step": round(e2e_ms_max / args.steps, 3),
                    "how": "PanguModel.forward on pinned host inputs; H2D/D2H double-buffered on side streams", "host": numa},
            "gpu_launches": launches,
            "roofline": roofline}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# training workload (BASELINE.json configs[3]: finetune_fully forward+backward, batch 1 per GPU, DDP gradient mean)
# ----------------------------------------------------------------------------------------------
def run_train_arm(args):
    import torch
    import torch.distributed as dist
    import pangu_pytorch_b200 as pb
    from pangu_pytorch_b200 import ops, training
    from pangu_pytorch_b200.dist import GradReducer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pb.set_operand_dtype(args.operands)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(0)
    model = pb.PanguModel(device=dev).to(dev).train()
    lora_mode = args.workload == "lora"
    if lora_mode:          # finetune/lora_tune.py:124-139 (lora_dropout 0 on this path)
        from pangu_pytorch_b200 import lora
        lora.add_lora(model, r=16, lora_alpha=16.0, lora_dropout=0.0)
        model.to(dev).train()
    if world > 1:
        model.grad_reducer = GradReducer()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=5e-6, weight_decay=3e-6,
                           fused=True)                                                   # finetune/finetune_fully.py:119
    g = torch.Generator(device=dev).manual_seed(1)
    maps = torch.randn(1, 3, 724, LON, device=dev, generator=g)
    const_h = torch.randn(1, 1, 1, 13, LAT, LON, device=dev, generator=g)
    stats = [torch.zeros(4, device=dev), torch.ones(4, device=dev),
             torch.zeros(13, 1, 1, 5, device=dev), torch.ones(13, 1, 1, 5, device=dev)]
    gk = torch.Generator(device=dev).manual_seed(100 + rank)
    up = torch.randn(1, 5, 13, LAT, LON, device=dev, generator=gk)
    sf = torch.randn(1, 4, LAT, LON, device=dev, generator=gk)
    tu = torch.randn(1, 5, 13, LAT, LON, device=dev, generator=gk)
    ts = torch.randn(1, 4, LAT, LON, device=dev, generator=gk)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = training.train_step(model, up, sf, stats, maps, const_h, tu, ts)
        opt.step()
        return loss

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = ops.launches() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    times = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_max = float(times[0])
    if rank == 0:
        pk = peaks()
        per = ms_max / args.steps
        tfl = 3 * FLOPS_PER_STEP / (per * 1e-3) / 1e12
        line = {"metric": ("lora_tune" if lora_mode else "finetune") + " steps/s @0.25deg (fwd + weighted-L1 + bwd + grad mean + Adam)", "value": round(world * args.steps / (ms_max / 1e3), 3),
                "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": round(per, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.operands, "data": "synthetic",
                "config": {"workload": ("lora_tune step (r=16 adapters on 67 nn.Linear, output convs trained in full)" if lora_mode else "finetune_fully step")
                                       + ", PanguModel 0.25deg, batch 1 per GPU, random-init weights",
                           "parallelism": f"data parallel over {world} GPU(s); gradient mean (1.1 GB fp32) all-reduced per block "
                                          "on a side stream under the backward", "optimizer": "torch.optim.Adam(fused), as the reference"},
                "clocks": clocks, "gpu_launches": launches, "loss": float(loss),
                "roofline": {"bound": "tensor", "achieved": round(tfl, 1), "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                             "frac": round(tfl / pk["tflops_sustained"], 4), "algorithmic_flops_per_step": 3 * FLOPS_PER_STEP,
                             "peak_source": pk["source"]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--operands", default=os.environ.get("PANGU_B200_OPERANDS", "bf16"), choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="forecast", choices=["forecast", "train", "lora"],
                    help="forecast: the BASELINE.json headline (default); train: finetune_fully step (configs[3]); "
                         "lora: lora_tune step (configs[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload in ("train", "lora"):
        run_train_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
