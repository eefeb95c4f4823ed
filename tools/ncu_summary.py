"""Turn ncu artefacts brought back in gpurun_out/ into small text summaries for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv  > profiles/r01_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/prof_X.ncu-rep    > profiles/r01_X.md
    python tools/ncu_summary.py traffic  gpurun_out/prof_X.ncu-rep profiles/r02_X.md lo  # records the capture's DRAM bytes in
                                                            # profiles/traffic.json under "<kernel>|lo" (bench.py reads it)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = [r for r in csv.DictReader(lines[start:]) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    for r in rows:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {len(rows)} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t:.3f} | {t / c:.4f} | {t / tot:.3f} |")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    print(f"# ncu --set full: {d.get('Kernel Name', ('?',))[0]}\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k][0]} | {d[k][1]} |")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"\nwarp-stall samples: {tot}; by reason: " +
          ", ".join(f"{h[6:]} {v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    print("\n| samples | executed | SASS | top stall |\n|---:|---:|---|---|")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
        st = max(((int(r[ix[h]] or 0), h[6:]) for h in stalls))
        print(f"| {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | `{r[ix['Source']].strip()[:70]}` | {st[1]} {st[0]} |")
    sass = " ".join(r[ix["Source"]] for r in data)
    print("\nBlackwell instructions present: " +
          ", ".join(f"{m}={'yes' if m in sass else 'no'}" for m in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "HMMA")))


def traffic(path, summary_md, tag):
    """Add dram__bytes_read + dram__bytes_write of this capture to profiles/traffic.json, keyed by kernel function
    (template arguments kept: CfgMLP1 and CfgLNRes384 are different kernels of the same engine)."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = sum(float(d[k][0]) * scale[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    name = re.sub(r"\(.*", "", d["Kernel Name"][0]).replace("void ", "").replace("pg::", "")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    db = json.load(open(out)) if os.path.exists(out) else {}
    db[f"{name}|{tag}"] = {"dram_bytes_per_launch": tot, "duration_us": float(d["gpu__time_duration.sum"][0]),
                "grid": int(float(d["launch__grid_size"][0])), "source": summary_md}
    with open(out, "w") as fh:
        json.dump(db, fh, indent=1, sort_keys=True)
    print(name, tot)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "traffic": lambda p: traffic(p, sys.argv[3], sys.argv[4])}[sys.argv[1]](sys.argv[2])
