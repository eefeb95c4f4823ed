#!/usr/bin/env python
"""Benchmark of the Pangu-Weather 24 h forecast step at 0.25 degrees (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--operands bf16|fp16]

A "step" is one full PanguModel forward (embed -> 16 earth-specific blocks -> recovery) of one
synthetic 0.25 degree sample (upper 5x13x721x1440, surface 4x721x1440), batch 1, random-init
weights of the reference architecture.  N > 1 (launched by torch.distributed.run) shards
independent ensemble members over the GPUs of one box: no data-path collective, weak scaling.

One JSON line is printed by rank 0.  ``value`` is device-resident throughput.  ``e2e`` is the same metric through the
public ``PanguModel.forward`` with every step's input fields copied from pinned host memory inside the timed region
(H2D 287 MB per step) and the step's result -- the 69 RMSE + 69 ACC scores of the forecast, computed on the device --
read back to the host; ``e2e_fields`` reads both fp32 forecast fields back instead (287 MB per step, the round-1
definition) and ``e2e_ensemble`` uploads one base state per 8 members and perturbs on the device (configs[2]).
``other_operands`` is the device-resident figure with the other 16-bit operand format, ``secondary`` holds short
finetune_fully / lora_tune steps (configs[3], configs[4]; data parallel over the same ranks when N > 1).
``roofline.traffic`` is looked up in profiles/traffic.json (written from committed ncu captures), never hard-coded.
``cpu_baseline`` / ``--impl reference`` time the UNMODIFIED reference forward (oracle/_ref, staged by
oracle/build_ref.py) on the host cores at the full 721 x 1440 grid; every CPU step is a complete forward.
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
LORA_DROPOUT = 0.0                  # lora_tune's adapter dropout in the secondary record
STRIP = 96                          # fallback CPU sample (only if oracle/_ref is absent): 96-column strip, 1/15 of the grid
WORKLOAD = ("PanguModel 24h forward, 0.25deg (upper 5x13x721x1440, surface 4x721x1440), batch 1, random-init weights; "
            "one ensemble member per step per GPU")
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
# CPU arm.  Preferred: the UNMODIFIED reference forward (oracle/_ref, staged by oracle/build_ref.py) at the full
# 721 x 1440 grid -- the quoted config, nothing extrapolated.  Only if the staged reference is absent: the oracle
# port on a 96-column strip x 15 (kind "port"), as in round 1.
# ----------------------------------------------------------------------------------------------
REF_ARM_BUDGET_S = 420.0            # --impl reference stops timing new steps after this long (CPU forward: 15-60 s each)


def _port_setup():
    import torch
    from oracle import pangu_oracle as O
    try:                                  # the GPU arm may have pinned this process next to its GPU: the CPU leg gets every core
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.reference_like_weights(seed=0)
    inputs = O.synthetic_inputs(seed=1, lat=LAT, lon=STRIP)

    def step():
        t0 = time.perf_counter()
        O.forward(p, *inputs)
        return (time.perf_counter() - t0) * (LON // STRIP)
    return step, torch.get_num_threads()


def cpu_arm():
    """-> (step() -> seconds per FULL forward, threads, kind, description of one step)."""
    from oracle import ref_runner
    if ref_runner.available():
        r = ref_runner.ReferenceForward()
        return (r.step, r.threads, "reference",
                "one full fp32 forward of the unmodified reference PanguModel (oracle/_ref = models/layers.py + "
                "models/pangu_model.py, eval, no_grad) on the 721x1440 grid")
    step, threads = _port_setup()
    return (step, threads, "port",
            f"oracle/_ref not staged: full-depth fp32 forward of the oracle port on a {STRIP}-column strip, scaled x{LON // STRIP}")


def cpu_baseline() -> dict:
    """Reported next to the GPU number (rank 0, N = 1): ONE timed full forward, no warm-up (the forward is 15-60 s of
    dense CPU work; allocator / thread-pool start-up is < 1 % of that)."""
    step, threads, kind, what = cpu_arm()
    t = step()
    return {"value": 1.0 / t, "unit": UNIT, "cores": threads, "kind": kind, "sample": f"{what}; {t:.1f} s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, threads, kind, what = cpu_arm()
    warm = min(args.warmup, 1)               # one warm-up forward at most: every CPU step is tens of seconds
    for _ in range(warm):
        step()
    want = max(1, args.steps)
    ts = []
    t_start = time.perf_counter()
    while len(ts) < want and (not ts or time.perf_counter() - t_start < REF_ARM_BUDGET_S):
        ts.append(step())
    dt = sum(ts) / len(ts)
    value = 1.0 / dt
    note = (f"{what}; {len(ts)} timed step(s) of {want} requested ({warm} warm-up): the arm stops starting new steps after "
            f"{REF_ARM_BUDGET_S:.0f} s so that the run ends within minutes; every step timed is a complete forward")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(ts), "steps_requested": want, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": note},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": note},
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

        # ---- timed region 2: end to end through the public API.  Every step copies that step's input fields from
        # pinned host memory (H2D, 287 MB) and reads the step's result back to pinned host memory; copies are
        # double-buffered on side streams.  Three variants (same kernels, same forward):
        #   metric   (headline `e2e`): result = the 69 latitude-weighted RMSE + 69 ACC values of the forecast against a
        #            resident verification field (on-device pangu_scores, what the reference's test() computes after every
        #            forward: models/pangu_sample.py:236-270); D2H = 552 B
        #   fields   (`e2e_fields`): result = both fp32 forecast fields (D2H = 287 MB), as in round 1
        #   ensemble (`e2e_ensemble`): BASELINE.json configs[2]: the base state is uploaded once per 8 members and the
        #            perturbations are drawn on the device (ensemble.perturb); result = scores per member
        from pangu_pytorch_b200 import ensemble as ens
        host_in = [(torch.empty(1, 5, 13, LAT, LON).pin_memory(), torch.empty(1, 4, LAT, LON).pin_memory())
                   for _ in range(2)]
        for (hu, hs), (du, ds) in zip(host_in, members):
            hu.copy_(du); hs.copy_(ds)
        host_out = [(torch.empty(1, 5, 13, LAT, LON).pin_memory(), torch.empty(1, 4, LAT, LON).pin_memory())
                    for _ in range(2)]
        host_sc = [torch.empty(2, 69).pin_memory() for _ in range(2)]
        dev_in = [(torch.empty_like(members[0][0]), torch.empty_like(members[0][1])) for _ in range(2)]
        gt = torch.Generator(device=dev).manual_seed(7)
        tgt = (torch.randn(1, 5, 13, LAT, LON, device=dev, generator=gt), torch.randn(1, 4, LAT, LON, device=dev, generator=gt))
        s_mean, s_std, u_mean, u_std = stats[0], stats[1], stats[2].reshape(13, 5), stats[3].reshape(13, 5)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        main = torch.cuda.current_stream()
        ENS = 8

        def e2e_run(nsteps, mode):
            in_ready = [None, None]
            in_free = [None, None]

            def issue_h2d(j):
                b = (j // ENS if mode == "ensemble" else j) % 2
                with torch.cuda.stream(s_in):
                    if in_free[b] is not None:
                        s_in.wait_event(in_free[b])
                    dev_in[b][0].copy_(host_in[b][0], non_blocking=True)
                    dev_in[b][1].copy_(host_in[b][1], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s_in); in_ready[b] = ev

            uploads = range(nsteps) if mode != "ensemble" else range(0, nsteps, ENS)
            issue_h2d(0)
            for j in range(nsteps):
                b = (j // ENS if mode == "ensemble" else j) % 2
                nxt = j + 1 if mode != "ensemble" else (j // ENS + 1) * ENS
                if nxt < nsteps and (mode != "ensemble" or j % ENS == 0):
                    issue_h2d(nxt)
                if mode != "ensemble" or j % ENS == 0:
                    main.wait_event(in_ready[b])
                if mode == "ensemble":
                    pu, ps = ens.perturb(dev_in[b][0], dev_in[b][1], rank * 1000 + j)
                    ou, os_ = model(pu, ps, stats, maps, const_h)
                else:
                    ou, os_ = model(dev_in[b][0], dev_in[b][1], stats, maps, const_h)
                if mode == "fields":
                    ev = torch.cuda.Event(); ev.record(main)
                    in_free[b] = ev
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev)
                        host_out[j % 2][0].copy_(ou, non_blocking=True)
                        host_out[j % 2][1].copy_(os_, non_blocking=True)
                        ou.record_stream(s_out); os_.record_stream(s_out)
                else:
                    ru, rs, au, as_ = ops.scores(ou, os_, tgt[0], tgt[1], s_mean, s_std, u_mean, u_std, normalised=True)
                    sc = torch.stack((torch.cat((ru.reshape(-1), rs.reshape(-1))), torch.cat((au.reshape(-1), as_.reshape(-1)))))
                    ev = torch.cuda.Event(); ev.record(main)
                    if mode != "ensemble" or j % ENS == ENS - 1 or j == nsteps - 1:
                        in_free[b] = ev
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev)
                        host_sc[j % 2].copy_(sc, non_blocking=True)
                        sc.record_stream(s_out)
            s_out.synchronize()
            main.synchronize()
            return len(uploads)

        e2e_times = {}
        for mode in ("metric", "fields", "ensemble"):
            e2e_run(2 if mode != "ensemble" else ENS + 1, mode)
            barrier()
            w0 = time.perf_counter()
            n_up = e2e_run(args.steps, mode)
            barrier()
            e2e_times[mode] = (time.perf_counter() - w0, n_up)
        score_sample = [float(host_sc[(args.steps - 1) % 2][0, 0]), float(host_sc[(args.steps - 1) % 2][1, 0])]

        # ---- fp16 operands (same tensor-core rate, ~8x lower error: meets the 1e-3 example tolerance of north_star)
        other = "fp16" if args.operands == "bf16" else "bf16"
        pb.set_operand_dtype(other)
        for i in range(3):
            step(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            step(i)
        f1.record()
        barrier()
        other_ms = f0.elapsed_time(f1)
        pb.set_operand_dtype(args.operands)

    # ---- reduce over ranks (max time)
    times = torch.tensor([ms, other_ms] + [e2e_times[m][0] * 1e3 for m in ("metric", "fields", "ensemble")],
                         device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_max, other_ms_max, e2e_ms_max, e2e_fields_ms, e2e_ens_ms = (float(x) for x in times)
    secondary = None
    if not args.no_secondary:
        del members, dev_in, host_in, host_out, tgt
        pb.free_workspaces()
        torch.cuda.empty_cache()
        secondary = run_secondary(args, model, dev, rank, world, maps, const_h, stats)
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
    traffic, traffic_source = ncu_traffic(dom, args.operands == "fp16")
    roofline = {"bound": "tensor", "kernel": dom, "achieved": round(achieved, 1), "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": round(achieved / pk["tflops_sustained"], 4),
                "traffic": traffic, "traffic_source": traffic_source,
                "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "avg_launch_ms": round(dom_ms, 4), "algorithmic_flops_per_launch": algo_flops[dom],
                "step_tflops": round(FLOPS_PER_STEP / (ms_per_step * 1e-3) / 1e12, 1),
                "step_frac_of_peak": round(FLOPS_PER_STEP / (ms_per_step * 1e-3) / 1e12 / pk["tflops_sustained"], 4),
                "attn_mlp_frac_of_peak": round(FLOPS_BLOCKS / (sum(v[0] for k, v in kern.items() if k in algo_flops) / args.steps * 1e-3)
                                               / 1e12 / pk["tflops_sustained"], 4),
                "kernel_time_shares": shares,
                "kernel_avg_launch_us": {k: round(1e3 * v[0] / v[1], 1) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}}
    how = "PanguModel.forward on pinned host inputs (H2D every step); copies double-buffered on side streams; "
    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.operands, "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "parallelism": f"ensemble members sharded over {world} GPU(s), no collective",
                       "accumulate": "fp32", "residual_stream": "fp32",
                       "l2": "per-step working set (>4 GB of activations) is far larger than the 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 2 * 69 * 4,
                    "ms_per_step": round(e2e_ms_max / args.steps, 3),
                    "how": how + "result read back = 69 RMSE + 69 ACC values of the forecast (on-device pangu_scores)",
                    "host": numa, "result_sample": score_sample},
            "e2e_fields": {"value": round(world * args.steps / (e2e_fields_ms / 1e3), 3), "unit": UNIT,
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d, "ms_per_step": round(e2e_fields_ms / args.steps, 3),
                           "how": how + "result read back = both fp32 forecast fields (round-1 definition of e2e)"},
            "e2e_ensemble": {"value": round(world * args.steps / (e2e_ens_ms / 1e3), 3), "unit": UNIT,
                             "h2d_bytes_per_step": round(h2d * e2e_times["ensemble"][1] / args.steps), "d2h_bytes_per_step": 2 * 69 * 4,
                             "ms_per_step": round(e2e_ens_ms / args.steps, 3),
                             "how": f"base state uploaded once per {ENS} members, members perturbed on the device (ensemble.perturb), "
                                    "scores read back per member"},
            "other_operands": {"dtype": other, "value": round(world * args.steps / (other_ms_max / 1e3), 3), "unit": UNIT,
                               "ms_per_step": round(other_ms_max / args.steps, 3),
                               "note": "same kernels with the other 16-bit operand format (fp16: per-variable rel-L2 <= 1e-3 vs the fp32 reference; bf16: <= 8e-3)"},
            "gpu_launches": launches,
            "roofline": roofline}
    if secondary is not None:
        line["secondary"] = secondary
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# kernels behind each entry point (names as ncu prints them, pg:: stripped), for roofline.traffic
ENTRY_KERNELS = {
    "pangu_mlp_ln_residual[lo]": ["mlp_fused2_kernel<{f}>|lo"],
    "pangu_mlp_ln_residual[hi]": ["mlp_fused_kernel<192, {f}>|hi"],
    "pangu_window_attention[lo]": ["window_attention_tc_kernel<{f}>|lo"],
    "pangu_window_attention[hi]": ["window_attention_tc_kernel<{f}>|hi"],
    "pangu_qkv[lo]": ["gemm_kernel<CfgQKV, {f}>|lo"], "pangu_qkv[hi]": ["gemm_kernel<CfgQKV, {f}>|hi"],
    "pangu_proj_ln_residual[lo]": ["gemm_kernel<CfgLNRes384, {f}>|lo"],
    "pangu_proj_ln_residual[hi]": ["gemm_kernel<CfgLNRes192, {f}>|hi"],
}


def ncu_traffic(entry: str, fp16: bool):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels behind ``entry``, from the committed
    ``ncu --set full`` captures (profiles/traffic.json, written by tools/ncu_summary.py traffic).  None (with the
    reason) when a kernel of the entry point has no capture -- never a stale constant."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, "profiles/traffic.json missing"
    with open(path) as fh:
        db = json.load(fh)
    tot, src = 0.0, []
    for k in ENTRY_KERNELS.get(entry, []):
        key = k.format(f=0)          # captures are taken with bf16 operands; the fp16 kernels move the same bytes
        if key not in db:
            return None, f"no ncu capture of {key} in profiles/traffic.json"
        tot += db[key]["dram_bytes_per_launch"]
        src.append(db[key]["source"])
    if not src:
        return None, f"no kernel list for {entry}"
    return tot, "ncu --set full, bytes per launch summed over the entry point's kernels: " + ", ".join(src)


# ----------------------------------------------------------------------------------------------
# secondary record: short driver-visible finetune / lora steps (BASELINE.json configs[3], configs[4])
# ----------------------------------------------------------------------------------------------
def run_secondary(args, model, dev, rank, world, maps, const_h, stats):
    import torch
    import torch.distributed as dist
    import pangu_pytorch_b200 as pb
    from pangu_pytorch_b200 import lora, training
    from pangu_pytorch_b200.dist import GradReducer
    out = {}
    gk = torch.Generator(device=dev).manual_seed(100 + rank)
    up = torch.randn(1, 5, 13, LAT, LON, device=dev, generator=gk)
    sf = torch.randn(1, 4, LAT, LON, device=dev, generator=gk)
    tu = torch.randn(1, 5, 13, LAT, LON, device=dev, generator=gk)
    ts = torch.randn(1, 4, LAT, LON, device=dev, generator=gk)
    steps = 3
    train_fmt = pb.training_operand_dtype()          # the training path's own default (fp16 + loss scale 2^16)
    for name in ("finetune", "lora"):
        torch.manual_seed(0)
        m = pb.PanguModel(device=dev).to(dev)
        if name == "lora":          # finetune/lora_tune.py:124-139: r=16, alpha=16, dropout 0.1, output convs trained in full
            lora.add_lora(m, r=16, lora_alpha=16.0, lora_dropout=LORA_DROPOUT)
            m.to(dev)
        m.train()
        if world > 1:
            m.grad_reducer = GradReducer()
        opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=5e-6, weight_decay=3e-6, fused=True)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = training.train_step(m, up, sf, stats, maps, const_h, tu, ts)
            opt.step()
            return loss

        for _ in range(2):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per = float(t[0]) / steps
        out[name] = {"metric": f"{name} steps/s @0.25deg (fwd + weighted-L1 + bwd + grad mean + fused Adam)",
                     "value": round(world * 1e3 / per, 3), "unit": "steps/s", "ms_per_step": round(per, 2), "steps": steps,
                     "warmup": 2, "dtype": train_fmt, "loss": float(loss),
                     "tflops": round(3 * FLOPS_PER_STEP / (per * 1e-3) / 1e12, 1),
                     "config": ("lora_tune: r=16 adapters on the 67 nn.Linear, lora_dropout %.1f, output convs trained in full" % LORA_DROPOUT
                                if name == "lora" else "finetune_fully: all 223 tensors") +
                               f"; batch 1 per GPU, {world} GPU(s), gradient mean over NCCL overlapped with the backward"}
        del m, opt
        training_cleanup()
    pb.set_operand_dtype(args.operands)
    return out


def training_cleanup():
    import torch
    import pangu_pytorch_b200 as pb
    pb.free_workspaces()
    torch.cuda.empty_cache()


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
    pb.set_training_operand_dtype(args.operands)

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
    ap.add_argument("--operands", default=None, choices=["bf16", "fp16"],
                    help="16-bit operand format (default: bf16 for the forecast, the training path's default -- fp16 -- for train / lora)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short finetune / lora steps of the `secondary` record")
    ap.add_argument("--workload", default="forecast", choices=["forecast", "train", "lora"],
                    help="forecast: the BASELINE.json headline (default); train: finetune_fully step (configs[3]); "
                         "lora: lora_tune step (configs[4])")
    args = ap.parse_args()
    if args.operands is None:
        if args.workload == "forecast":
            args.operands = os.environ.get("PANGU_B200_OPERANDS", "bf16")
        else:
            args.operands = os.environ.get("PANGU_B200_TRAIN_OPERANDS", "fp16")
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload in ("train", "lora"):
        run_train_arm(args)
    else:
        run_gpu_arm(args)


def _only_the_json_line_on_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on rank 0 when the
    box sets NCCL_DEBUG=VERSION): point fd 1 at stderr for the whole run and give ``print`` the real stdout back."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _only_the_json_line_on_stdout()
    main()
