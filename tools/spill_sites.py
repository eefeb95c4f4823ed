"""Where does a kernel spill?  Maps STL/LDL instructions of one kernel of libpangu_b200.so to CUDA source lines
(needs -lineinfo).   python tools/spill_sites.py window_attention_tc_kernelILb0"""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pangu_pytorch_b200", "libpangu_b200.so")
pat = sys.argv[1]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
inside, cur, c = False, None, collections.Counter()
for l in txt.splitlines():
    if l.startswith("//---") and ".text." in l:
        inside = pat in l
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    if re.search(r"\b(STL|LDL)\b", l):
        c[(cur, "STL" if "STL" in l else "LDL")] += 1
for k, v in sorted(c.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(k, v)
