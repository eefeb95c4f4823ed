"""Tabulate a tools/attn_only.py trace (ATTN_TRACE=1): per window, the softmax pair's phases relative to warp A's entry.
    python tools/trace_table.py gpurun_out/xxx_attn_trace_lo.log [first] [last]"""
import re, sys
txt = open(sys.argv[1]).read()
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 34)
roles, cur = {}, None
for l in txt.splitlines():
    if l.startswith('   win'):
        m = re.match(r'\s+win\s+(\d+): \[(.*)\]', l)
        roles[cur][int(m.group(1))] = [int(x) for x in m.group(2).split(',')]
    elif l[:4] in ('TMA ', 'MMA ', 'TAIL', 'SOFT', 'SEG ', 'EXP ', 'MAX ') or l.strip() == '-':
        cur = l.strip(); roles[cur] = {}
names = list(roles)
tma, mma, tail, s3, s4, s5, seg = [roles[n] for n in names[:7]]
print("win | EXP warp: S+max ready, exp done, P published, next enter | MAX warp (rel. to EXP enter): enter, S ready, max posted, epilogue done | MMA saw P | tail")
for w in range(lo, hi):
    e = s3[w][0]
    print(w, [x - e for x in s3[w][1:]] + [s3[w + 1][0] - e], [x - e for x in s5[w]], mma[w][2] - e, '| tail', tail[w][1] - tail[w][0])
print("segments [start, bias staged, all done]:", [seg[k][:3] for k in sorted(seg) if seg[k][0] >= 0][:6])
