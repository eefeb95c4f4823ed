"""Top CUDA source lines by warp-stall samples of an ncu capture (compiled with -lineinfo, --import-source on).
    python tools/ncu_lines.py gpurun_out/prof_X.ncu-rep [min_samples]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hdr = next(r for r in rows[:6] if "# Samples" in r)
si = hdr.index("# Samples")
out = []
for r in rows:
    if len(r) <= si or r is hdr:
        continue
    try:
        n = int(r[si] or 0)
    except ValueError:
        continue
    if n >= thr and r[3].strip() == "-":          # aggregated CUDA-source row (SASS column empty)
        out.append((n, r[0], r[1].strip()[:150], r[4][:80]))
out.sort(reverse=True)
for o in out[:30]:
    print(o)
