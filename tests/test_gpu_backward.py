"""Kernel-level parity of the backward-pass kernels (C ABI section "Backward pass") against fp64
autograd / matmul restatements of the same op.  Whole-model gradient parity against the oracle's
autograd lives in tests/test_gpu_training.py.

Tolerances: the kernels consume 16-bit operands that are generated already rounded, so against an
fp64 evaluation of the same rounded operands only fp32 accumulation (and, for 16-bit outputs, the
final rounding: bf16 2^-9, fp16 2^-12 relative) remains.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import pangu_oracle as O
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FORMATS = ["bf16", "fp16"]
OUT16_TOL = {"bf16": 6e-3, "fp16": 1e-3}


def _fmt(fmt):
    import pangu_pytorch_b200 as pb
    pb.set_operand_dtype(fmt)
    pb.free_workspaces()
    return fmt == "fp16"


def _r16(t, fp16):
    return t.to(torch.float16 if fp16 else torch.bfloat16)


@pytest.mark.parametrize("fmt", FORMATS)
def test_cast16_t_exact(fmt):
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    w = torch.randn(160, 384, generator=torch.Generator().manual_seed(0)).to(DEV)
    out = ops.cast16_t(w, fp16, rows_pad=192)
    ref = torch.zeros(384, 192, dtype=out.dtype, device=DEV)
    ref[:, :160] = _r16(w, fp16).t()
    assert torch.equal(out, ref)


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("M,N,K,nv,kv", [(1000, 768, 192, None, None), (4133, 192, 768, None, None),
                                         (20000, 576, 192, None, None), (7777, 384, 1536, None, None),
                                         (3000, 192, 192, 160, None), (3000, 192, 128, None, 112),
                                         (130, 64, 384, None, None)])
def test_wgrad(fmt, M, N, K, nv, kv):
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    g = torch.Generator().manual_seed(M + N + K)
    dy = _r16(torch.randn(M, N, generator=g), fp16).to(DEV)
    x = _r16(torch.randn(M, K, generator=g), fp16).to(DEV)
    n_out, k_out = nv or N, kv or K
    dw0 = torch.randn(n_out, k_out + 8, generator=g).to(DEV)       # wider buffer + column offset: ldw / k_off path
    dw = dw0.clone()
    ops.wgrad(dy, x, dw, fp16, n_valid=nv, k_valid=kv, k_off=4, alpha=0.5)
    torch.cuda.synchronize()
    ref = dw0.double()
    ref[:, 4:4 + k_out] += 0.5 * (dy.double().t() @ x.double())[:n_out, :k_out]
    assert rel_l2(dw, ref) < 1e-5
    assert torch.equal(dw[:, :4], dw0[:, :4]) and torch.equal(dw[:, 4 + k_out:], dw0[:, 4 + k_out:])


@pytest.mark.parametrize("fmt", FORMATS)
def test_colsum16(fmt):
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    g = torch.Generator().manual_seed(3)
    for M, N, nv in [(5000, 192, None), (3333, 1536, None), (999, 192, 160), (4000, 64, None)]:
        src = _r16(torch.randn(M, N, generator=g), fp16).to(DEV)
        out0 = torch.randn(nv or N, generator=g).to(DEV)
        out = out0.clone()
        ops.colsum16(src, out, fp16, n_valid=nv, alpha=2.0)
        ref = out0.double() + 2.0 * src.double().sum(0)[: nv or N]
        assert rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("fmt", FORMATS)
def test_dgrad_plain_and_16bit(fmt):
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    g = torch.Generator().manual_seed(11)
    M, N, K = 3001, 384, 768
    a = _r16(torch.randn(M, K, generator=g), fp16).to(DEV)
    wt = _r16(torch.randn(N, K, generator=g) * 0.05, fp16).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = a.double() @ wt.double().t()
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.dgrad(a, wt, 0, fp16, out32=out)
    assert rel_l2(out, ref) < 2e-6
    out = res.clone()
    ops.dgrad(a, wt, 0, fp16, out32=out, resid32=out)                  # in-place accumulate
    assert rel_l2(out, ref + res.double()) < 2e-6
    wt2 = _r16(torch.randn(768, K, generator=g) * 0.05, fp16).to(DEV)
    b = torch.randn(768, generator=g).to(DEV)
    out16 = torch.empty(M, 768, dtype=a.dtype, device=DEV)
    ops.dgrad(a, wt2, 1, fp16, out16=out16, bias=b)
    assert rel_l2(out16.float(), a.double() @ wt2.double().t() + b.double()) < OUT16_TOL[fmt]


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("H,C", [(181, 192), (91, 384)])
@pytest.mark.parametrize("roll", [False, True])
def test_dgrad_window_maps(fmt, H, C, roll):
    """kind 2 (window-order rows -> token scatter-add) and kind 3 (token rows -> window-order scatter)."""
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    Z, W = 8, 24
    T = Z * H * W
    src = O.window_source_index(Z, H, W, roll).reshape(-1)
    Tp = src.numel()
    real = src >= 0
    g = torch.Generator().manual_seed(H + int(roll))
    K = 192
    wt = _r16(torch.randn(C, K, generator=g) * 0.05, fp16).to(DEV)
    # kind 2
    a = _r16(torch.randn(Tp, K, generator=g), fp16).to(DEV)
    res = torch.randn(T, C, generator=g).to(DEV)
    out = res.clone()
    ops.dgrad(a, wt, 2, fp16, out32=out, resid32=out, grid=(Z, H, W), roll=roll)
    full = (a.double() @ wt.double().t()).cpu()
    ref = res.double().cpu()
    ref[src[real]] += full[real]
    assert rel_l2(out, ref) < 2e-6
    # kind 3
    a = _r16(torch.randn(T, K, generator=g), fp16).to(DEV)
    out16 = torch.zeros(Tp, C, dtype=a.dtype, device=DEV)
    ops.dgrad(a, wt, 3, fp16, out16=out16, grid=(Z, H, W), roll=roll)
    full = (a.double() @ wt.double().t()).cpu()
    ref = torch.zeros(Tp, C, dtype=torch.float64)
    ref[real] = full[src[real]]
    assert rel_l2(out16.float(), ref) < OUT16_TOL[fmt]
    assert (out16[~real.to(DEV)] == 0).all()


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("C", [192, 384])
def test_layernorm_bwd_plain(fmt, C):
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    g = torch.Generator().manual_seed(C)
    rows, scale = 5003, 0.8
    y = (torch.randn(rows, C, generator=g) * 2 + 0.5).to(DEV)
    go = torch.randn(rows, C, generator=g).to(DEV)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = torch.randn(C, generator=g).to(DEV)
    yd = y.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    (scale * F.layer_norm(yd, (C,), gd, bd, 1e-5) * go.double()).sum().backward()
    dx16 = torch.empty(rows, C, dtype=torch.float16 if fp16 else torch.bfloat16, device=DEV)
    dg0, db0 = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    dg, db = dg0.clone(), db0.clone()
    dbias = torch.zeros(C, device=DEV)
    ops.layernorm_bwd(y, go, gamma, dg, db, rows, C, 0, fp16, dx16=dx16, scale=scale, dbias=dbias, palpha=0.5)
    dg, db = dg0 + 2 * (dg - dg0), db0 + 2 * (db - db0)           # undo palpha = 0.5
    assert rel_l2(dx16.float(), yd.grad) < OUT16_TOL[fmt]
    assert rel_l2(2 * dbias, yd.grad.sum(0)) < 1e-3              # fused bias gradient of the producing linear
    assert rel_l2(dg - dg0, gd.grad) < 1e-4
    assert rel_l2(db - db0, bd.grad) < 1e-4


@pytest.mark.parametrize("fmt", FORMATS)
def test_layernorm_bwd_upsample_rows(fmt):
    """mode 1: rows are high-res tokens whose pre-norm values sit pixel-shuffled in the [T2, 768] linear1 output."""
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    Z, H, W = 8, 181, 24
    H2, W2, C = 91, 12, 192
    g = torch.Generator().manual_seed(1)
    u1 = torch.randn(Z * H2 * W2, 4 * C, generator=g).to(DEV)
    go = torch.randn(Z * H * W, C, generator=g).to(DEV)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    ud = u1.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    v = ud.view(Z, H2, W2, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(Z, 2 * H2, 2 * W2, C)[:, :H].reshape(-1, C)
    (F.layer_norm(v, (C,), gd, torch.zeros(C, dtype=torch.float64, device=DEV), 1e-5) * go.double()).sum().backward()
    du = torch.zeros(Z * H2 * W2, 4 * C, dtype=torch.float16 if fp16 else torch.bfloat16, device=DEV)
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.layernorm_bwd(u1, go, gamma, dg, db, Z * H * W, C, 1, fp16, dx16=du, grid=(Z, H, W))
    assert rel_l2(du.float(), ud.grad) < OUT16_TOL[fmt]
    assert rel_l2(dg, gd.grad) < 1e-4
    assert rel_l2(db, go.double().sum(0)) < 1e-4


def test_layernorm_bwd_downsample_rows():
    """mode 2: rows are low-res tokens; the 2x2 merge (+ zero pad row) is recomputed from the high-res stream."""
    from pangu_pytorch_b200 import ops
    _fmt("bf16")
    Z, H, W, C = 8, 181, 24, 192
    H2, W2 = 91, 12
    g = torch.Generator().manual_seed(2)
    x = torch.randn(Z * H * W, C, generator=g).to(DEV)
    go = torch.randn(Z * H2 * W2, 4 * C, generator=g).to(DEV)
    gamma = (1 + 0.2 * torch.randn(4 * C, generator=g)).to(DEV)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    m = F.pad(xd.view(Z, H, W, C), (0, 0, 0, 0, 0, 1)).view(Z, H2, 2, W2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, 4 * C)
    (F.layer_norm(m, (4 * C,), gd, torch.zeros(4 * C, dtype=torch.float64, device=DEV), 1e-5) * go.double()).sum().backward()
    base = torch.randn(Z * H * W, C, generator=g).to(DEV)
    dx = base.clone()
    dg, db = torch.zeros(4 * C, device=DEV), torch.zeros(4 * C, device=DEV)
    ops.layernorm_bwd(x, go, gamma, dg, db, Z * H2 * W2, 4 * C, 2, False, dx32=dx, grid=(Z, H, W))
    assert rel_l2(dx - base, xd.grad) < 1e-5
    assert rel_l2(dg, gd.grad) < 1e-4


@pytest.mark.parametrize("fmt", FORMATS)
def test_gelu_bwd(fmt):
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    g = torch.Generator().manual_seed(4)
    pre = _r16(torch.randn(1000, 768, generator=g) * 2, fp16).to(DEV)
    dh = _r16(torch.randn(1000, 768, generator=g), fp16).to(DEV)
    pd = pre.double().requires_grad_(True)
    (F.gelu(pd) * dh.double()).sum().backward()
    db = torch.zeros(768, device=DEV)
    ops.gelu_bwd(dh, pre, fp16, dbias=db, alpha=2.0)
    assert rel_l2(dh.float(), pd.grad) < OUT16_TOL[fmt]
    assert rel_l2(db, 2.0 * pd.grad.sum(0)) < 1e-3               # fused bias gradient (column sums)
    # ragged row count, wide rows, no bias output
    pre2 = _r16(torch.randn(333, 1536, generator=g), fp16).to(DEV)
    dh2 = _r16(torch.randn(333, 1536, generator=g), fp16).to(DEV)
    p2 = pre2.double().requires_grad_(True)
    (F.gelu(p2) * dh2.double()).sum().backward()
    ops.gelu_bwd(dh2, pre2, fp16)
    assert rel_l2(dh2.float(), p2.grad) < OUT16_TOL[fmt]


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("H,C,heads", [(181, 192, 6), (91, 384, 12)])
@pytest.mark.parametrize("roll", [False, True])
def test_window_attention_bwd(fmt, H, C, heads, roll):
    """dq/dk/dv and d earth_specific_bias against fp64 autograd of the oracle's attention math on the
    same 16-bit q/k/v (q pre-scaled, head-major planes as pangu_qkv writes them)."""
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    Z, W = 8, 24
    nLon, types = W // 12, 4 * ((H + 5) // 6)
    Tp = nLon * types * 144
    Tpp = (Tp + 127) // 128 * 128
    g = torch.Generator().manual_seed(H + int(roll))
    h16 = torch.float16 if fp16 else torch.bfloat16
    qkv = torch.zeros(3 * heads, Tpp, 32, dtype=h16)
    qkv[:, :Tp] = _r16(torch.randn(3 * heads, Tp, 32, generator=g), fp16)
    bias = torch.randn(1, types, heads, 144, 144, generator=g)
    datt = _r16(torch.randn(Tp, C, generator=g), fp16)
    scale = 32 ** -0.5
    # fp64 reference: q_raw is the un-scaled projection (stored q = q_raw * scale, rounded)
    qs = qkv[:heads, :Tp].double().view(heads, nLon, types, 144, 32).permute(1, 2, 0, 3, 4)
    q_raw = (qs / scale).clone().requires_grad_(True)
    k = qkv[heads:2 * heads, :Tp].double().view(heads, nLon, types, 144, 32).permute(1, 2, 0, 3, 4).clone().requires_grad_(True)
    v = qkv[2 * heads:, :Tp].double().view(heads, nLon, types, 144, 32).permute(1, 2, 0, 3, 4).clone().requires_grad_(True)
    bd = bias.double().clone().requires_grad_(True)
    s = (q_raw * scale) @ k.transpose(-2, -1) + bd
    if roll:
        s = s + O.shift_mask(Z, H).double().view(1, types, 1, 144, 144)
    o = (torch.softmax(s, -1) @ v).permute(0, 1, 3, 2, 4).reshape(Tp, C)
    (o * datt.double()).sum().backward()
    dqkv = torch.full((Tp, 3 * C), float("nan"), dtype=h16, device=DEV)
    dbias = torch.zeros(types, heads, 144, 144, device=DEV)
    dbqkv = torch.zeros(3 * C, device=DEV)
    ops.window_attention_bwd(qkv.to(DEV), datt.to(DEV), bias.to(DEV), dqkv, dbias, Z, H, W, C, heads, roll, fp16, dbqkv=dbqkv)
    torch.cuda.synchronize()

    def planes(t):     # [nLon, types, heads, 144, 32] -> [Tp, C]
        return t.permute(0, 1, 3, 2, 4).reshape(Tp, C)
    tol = 2e-2 if not fp16 else 3e-3       # P and dS enter the second GEMMs rounded to 16 bits
    out = dqkv.float().cpu()
    assert torch.isfinite(out).all()
    assert rel_l2(out[:, :C], planes(q_raw.grad)) < tol
    assert rel_l2(out[:, C:2 * C], planes(k.grad)) < tol
    assert rel_l2(out[:, 2 * C:], planes(v.grad)) < tol
    assert rel_l2(dbias, bd.grad[0]) < tol
    ref_cols = torch.cat([planes(q_raw.grad).sum(0), planes(k.grad).sum(0), planes(v.grad).sum(0)])
    assert rel_l2(dbqkv, ref_cols) < tol                          # fused attention.linear1.bias gradient


@pytest.mark.parametrize("fmt", FORMATS)
def test_recover_grad_gather(fmt):
    from pangu_pytorch_b200 import ops
    fp16 = _fmt(fmt)
    lat, lon = 721, 96
    Hh, Ww = 181, 24
    g = torch.Generator().manual_seed(6)
    du = _r16(torch.randn(1, 5, 13, lat, lon, generator=g), fp16).float()
    ds = _r16(torch.randn(1, 4, lat, lon, generator=g), fp16).float()
    yu = torch.zeros(7 * Hh * Ww, 160, requires_grad=True)
    ys = torch.zeros(Hh * Ww, 64, requires_grad=True)
    fu = yu.view(7, Hh, Ww, 5, 2, 4, 4).permute(3, 0, 4, 1, 5, 2, 6).reshape(5, 14, 4 * Hh, 4 * Ww)[:, :13, :lat]
    fs = ys.view(Hh, Ww, 4, 4, 4).permute(2, 0, 3, 1, 4).reshape(4, 4 * Hh, 4 * Ww)[:, :lat]
    ((fu * du[0]).sum() + (fs * ds[0]).sum()).backward()
    h16 = torch.float16 if fp16 else torch.bfloat16
    dyu = torch.full((7 * Hh * Ww, 192), float("nan"), dtype=h16, device=DEV)
    dys = torch.full((Hh * Ww, 128), float("nan"), dtype=h16, device=DEV)
    ops.recover_grad_gather(du.to(DEV), ds.to(DEV), dyu, dys, lat, lon, fp16)
    assert torch.equal(dyu[:, :160].float().cpu(), yu.grad) and (dyu[:, 160:] == 0).all()
    assert torch.equal(dys[:, :64].float().cpu(), ys.grad) and (dys[:, 64:] == 0).all()
