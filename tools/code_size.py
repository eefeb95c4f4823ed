"""Instruction count of one kernel by CUDA source line range (needs -lineinfo): where does the code size go?
    python tools/code_size.py window_attention_tc_kernelILb0 [bucket]"""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pangu_pytorch_b200", "libpangu_b200.so")
pat = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 20
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
inside, cur, c, tot = False, None, collections.Counter(), 0
for l in txt.splitlines():
    if l.startswith("//---") and ".text." in l:
        inside = pat in l
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)) // bucket * bucket)
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        c[cur] += 1
        tot += 1
print("total instructions", tot, "=", tot * 16 // 1024, "KB")
for k, v in sorted(c.items(), key=lambda kv: (kv[0] or ("", 0))):
    if v >= 20:
        print(k, v)
