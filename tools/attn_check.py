"""Window-attention kernel vs a torch fp32 reference on random head-major qkv (debug aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pangu_pytorch_b200 import engine, ops

tag = sys.argv[1] if len(sys.argv) > 1 else "hi"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 24
Z, H, C, heads = (8, 181, 192, 6) if tag == "hi" else (8, 91, 384, 12)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ws = engine.workspace(dev, Z, H, W, C)
Tp, types, nlon = ws.Tp, ws.types, ws.nlon
ws.qkv.normal_()
ebias = torch.randn(1, types, heads, 144, 144, device=dev)
for roll in (0, 1):
    ws.att.zero_()
    ops.window_attention(ws.qkv, ebias, ws.att, Z, H, W, C, heads, roll, False, True)
    torch.cuda.synchronize()
    q, k, v = [ws.qkv[s * heads:(s + 1) * heads, :Tp].float().view(heads, nlon, types, 144, 32) for s in range(3)]
    S = torch.einsum("hltid,hltjd->hltij", q, k) + ebias[0].permute(1, 0, 2, 3)[:, None]
    if roll:
        nH = (H + 5) // 6
        kk = torch.arange(144, device=dev)
        zl, hl = kk // 72, (kk // 12) % 6
        m = torch.zeros(types // nH, nH, 144, 144, dtype=torch.bool, device=dev)
        m[-1] |= zl[:, None] != zl[None, :]
        m[:, -1] |= (hl[:, None] < 3) != (hl[None, :] < 3)
        S = S + torch.where(m.view(types, 144, 144), -100.0, 0.0)[None, None]
    P = torch.softmax(S, -1)
    ref = torch.einsum("hltij,hltjd->hltid", P, v)                     # [heads, nlon, types, 144, 32]
    got = ws.att.float().view(nlon, types, 144, heads, 32).permute(3, 0, 1, 2, 4)
    err = (got - ref).norm(dim=(-1, -2)) / ref.norm(dim=(-1, -2))       # per (head, lon, type)
    bad = (err > 0.02).nonzero()
    print(f"{tag} W={W} roll={roll}: overall rel {float((got - ref).norm() / ref.norm()):.3e}; bad (head,lon,type) units: {len(bad)} of {err.numel()}")
    for b in bad[:12]:
        h_, l_, t_ = [int(x) for x in b]
        d = (got[h_, l_, t_] - ref[h_, l_, t_]).norm(dim=-1) / ref[h_, l_, t_].norm(dim=-1)
        u = (t_ * heads + h_) * nlon + l_
        print(f"   unit u={u} head {h_} lon {l_} type {t_}: rows<128 err {float(d[:128].mean()):.3f} tail rows err {float(d[128:].mean()):.3f}")
