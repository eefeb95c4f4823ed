"""Training path of ``PanguModel``: forward that keeps the activations the backward needs, and the
hand-written backward pass (reference: ``loss.backward()`` through
models/layers.py, models/pangu_sample.py:52-71).

The reference wraps every block in ``torch.utils.checkpoint`` (models/layers.py:118-123) because a
0.25 degree sample does not fit a 2019-2023 GPU with autograd's saved tensors.  On a B200 the
per-block operands (window-ordered input, head-major q/k/v, merged attention output, the two Mlp
operands: 2.0 GB per 192-channel block, 1.0 GB per 384-channel block, 20.4 GB in total, all 16-bit)
simply stay resident in HBM in a ``Tape`` that is allocated once and reused by every step, so
nothing but three small GEMMs per block (the pre-LayerNorm / pre-GELU values) is recomputed.

Gradients are produced by the kernels declared under "Backward pass" in ``include/pangu_b200.h``:
tcgen05 dgrad / wgrad GEMMs, the window-attention backward, LayerNorm / GELU backward and bias
column sums.  torch is used for buffers, streams and the autograd hand-over only.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import engine, ops
from .engine import Weight16T
from .lora import LoraLinear

LNB_IDENT, LNB_UP, LNB_DOWN = 0, 1, 2
# Loss scale of the backward pass.  dL/d(output) of the weighted L1 loss is ~1e-8 at 0.25 degrees: below the fp16
# subnormal range, so with fp16 operands the seed is multiplied by 2^16 when it is patchified and every parameter
# gradient is multiplied by 2^-16 where it is accumulated (exact powers of two).  bf16 has the fp32 range: 1.
LOSS_SCALE = {False: 1.0, True: 65536.0}


class BlockTape:
    def __init__(self, ws: engine.GridWorkspace):
        dev, h = ws.x32.device, ops.dtype16(ws.fp16)
        self.xw = torch.zeros(ws.Tp, ws.C, dtype=h, device=dev)        # pad rows stay zero
        self.qkv = torch.empty_like(ws.qkv)
        self.att = torch.empty(ws.T, ws.C, dtype=h, device=dev)
        self.xmid16 = torch.empty(ws.T, ws.C, dtype=h, device=dev)
        self.hidden = torch.empty(ws.T, 4 * ws.C, dtype=h, device=dev)
        self.s1 = self.s2 = 1.0


class Scratch:
    """Backward work buffers of one resolution."""

    def __init__(self, ws: engine.GridWorkspace):
        dev, h, f = ws.x32.device, ops.dtype16(ws.fp16), torch.float32
        T, Tp, C = ws.T, ws.Tp, ws.C
        self.y32 = torch.empty(T, C, dtype=f, device=dev)
        self.tmp16 = torch.empty(T, C, dtype=h, device=dev)
        self.dy16 = torch.empty(T, C, dtype=h, device=dev)
        self.dh16 = torch.empty(T, 4 * C, dtype=h, device=dev)
        self.pre16 = torch.empty(T, 4 * C, dtype=h, device=dev)
        self.dattw = [torch.zeros(Tp, C, dtype=h, device=dev) for _ in range(2)]     # per roll state; pad rows stay zero
        self.dqkv = torch.empty(Tp, 3 * C, dtype=h, device=dev)
        self.g32 = torch.empty(T, C, dtype=f, device=dev)                            # gradient of the residual stream


class Tape:
    def __init__(self, model, hi: engine.GridWorkspace, lo: engine.GridWorkspace):
        dev, h, f = hi.x32.device, ops.dtype16(hi.fp16), torch.float32
        self.hi, self.lo = hi, lo
        self.fp16 = hi.fp16
        self.order = block_order(model, hi, lo)
        self.blocks = [BlockTape(ws) for (_, ws, _) in self.order]
        plane = hi.H * hi.W
        self.a_u = torch.empty(7 * plane, 192, dtype=h, device=dev)
        self.a_s = torch.empty(plane, 128, dtype=h, device=dev)
        self.skip16 = torch.empty(hi.T, hi.C, dtype=h, device=dev)
        self.x32_skip = torch.empty(hi.T, hi.C, dtype=f, device=dev)
        self.down_a = torch.empty(lo.T, 4 * hi.C, dtype=h, device=dev)
        self.lo_out16 = torch.empty(lo.T, lo.C, dtype=h, device=dev)
        self.up_a = torch.empty(hi.T, hi.C, dtype=h, device=dev)
        self.final16 = torch.empty(hi.T, hi.C, dtype=h, device=dev)
        self.scr = {id(hi): Scratch(hi), id(lo): Scratch(lo)}
        # up / down-sample and recovery work buffers
        self.u32 = torch.empty(lo.T, 4 * hi.C, dtype=f, device=dev)
        self.u16 = torch.empty(lo.T, 4 * hi.C, dtype=h, device=dev)
        self.du16 = torch.empty(lo.T, 4 * hi.C, dtype=h, device=dev)
        self.dyu = torch.empty(7 * plane, 192, dtype=h, device=dev)
        self.dys = torch.empty(plane, 128, dtype=h, device=dev)
        self.g_skip = torch.empty(hi.T, hi.C, dtype=f, device=dev)
        self.lat = self.lon = 0
        self.generation = 0          # bumped by every training-mode forward: a backward must match the forward that filled the tape


def block_order(model, hi, lo):
    """The 16 blocks in execution order as (block, workspace, roll) (models/layers.py:116-124)."""
    out = []
    for li, layer in enumerate(model.layers):
        ws = hi if li in (0, 3) else lo
        for i, blk in enumerate(layer.blocks):
            out.append((blk, ws, i % 2 == 1))
    return out


def _tape(model, hi, lo) -> Tape:
    key = (id(hi), id(lo), hi.fp16)
    t = getattr(model, "_tape", None)
    if t is None or t[0] != key:
        model._tape = (key, Tape(model, hi, lo))
    return model._tape[1]


def release_tape(model) -> None:
    """Free the saved-activation buffers (about 25 GB at 0.25 degrees)."""
    if hasattr(model, "_tape"):
        del model._tape


# ----------------------------------------------------------------------------------------------
# forward (training mode)
# ----------------------------------------------------------------------------------------------
def forward_train(model, input, input_surface, statistics, maps, const_h):
    """Same kernels as the inference forward; every per-block buffer is redirected into the tape."""
    dev = model._input_layer.conv.weight.device
    lat, lon = input_surface.shape[-2], input_surface.shape[-1]
    if input.shape[0] != 1:
        raise ValueError("batch size is 1 on this path, as in the reference (models/layers.py:219,227)")
    Z, H, W = 8, (lat + 3) // 4, lon // 4
    hi = engine.workspace(dev, Z, H, W, 192)
    lo = engine.workspace(dev, Z, (H + 1) // 2, W // 2, 384)
    tape = _tape(model, hi, lo)
    tape.lat, tape.lon = lat, lon
    tape.generation += 1
    fp16 = hi.fp16
    order, bt = tape.order, tape.blocks
    # what each block writes its 16-bit shadow to: the next block's window-ordered input, or a natural-order buffer
    natural_out = {1: tape.skip16, 13: tape.lo_out16, 15: tape.final16}
    saved = {id(ws): (ws.x16w[0], ws.x16w[1], ws.qkv, ws.att, ws.x16, ws.hidden) for ws in (hi, lo)}
    try:
        emb = model._input_layer
        s_mean, s_std = engine.f32(statistics[0], dev).reshape(4), engine.f32(statistics[1], dev).reshape(4)
        u_mean, u_std = engine.f32(statistics[2], dev).reshape(13, 5), engine.f32(statistics[3], dev).reshape(13, 5)
        hi.x16w[0] = bt[0].xw
        ops.patch_embed(engine.f32(input, dev), engine.f32(input_surface, dev), s_mean, s_std, u_mean, u_std,
                        engine.f32(maps, dev), engine.f32(const_h, dev), emb._w.get(emb.conv.weight), emb.conv.bias,
                        emb._ws.get(emb.conv_surface.weight), emb.conv_surface.bias, tape.a_u, tape.a_s, hi.x32,
                        hi.x16w[0], lat, lon, fp16)

        def run_block(b):
            blk, ws, roll = order[b]
            t = bt[b]
            ws.qkv, ws.att, ws.x16, ws.hidden = t.qkv, t.att, t.xmid16, t.hidden
            ws.x16w[int(roll)] = t.xw
            if b in natural_out:
                t.s1, t.s2 = blk._run(ws, roll, -1, natural_out[b])
            else:
                nroll = int(order[b + 1][2])
                ws.x16w[nroll] = bt[b + 1].xw
                t.s1, t.s2 = blk._run(ws, roll, nroll)

        run_block(0); run_block(1)
        tape.x32_skip.copy_(hi.x32)
        down = model.downsample
        lo.x16w[0] = bt[2].xw
        ops.downsample(hi.x32, down.norm.weight, down.norm.bias, down._w.get(down.linear.weight), tape.down_a, lo.x32,
                       lo.x16w[0], hi.Z, hi.H, hi.W, hi.C, fp16)
        for b in range(2, 14):
            run_block(b)
        up = model.upsample
        hi.x16w[0] = bt[14].xw
        ops.upsample(tape.lo_out16, up._w1.get(up.linear1.weight), up.norm.weight, up.norm.bias,
                     up._w2.get(up.linear2.weight), tape.up_a, hi.x32, hi.x16w[0], hi.Z, hi.H, hi.W, lo.C, hi.C, fp16)
        run_block(14); run_block(15)
        rec = model._output_layer
        out = torch.empty(1, 5, 13, lat, lon, dtype=torch.float32, device=dev)
        out_s = torch.empty(1, 4, lat, lon, dtype=torch.float32, device=dev)
        ops.patch_recover(tape.skip16, tape.final16, rec._w.get(rec.conv.weight), rec.conv.bias,
                          rec._ws.get(rec.conv_surface.weight), rec.conv_surface.bias, out, out_s,
                          hi.Z, hi.H, hi.W, hi.C, lat, lon, fp16)
    finally:
        for ws in (hi, lo):
            a, b, ws.qkv, ws.att, ws.x16, ws.hidden = saved[id(ws)]
            ws.x16w[0], ws.x16w[1] = a, b
    return out, out_s, tape


# ----------------------------------------------------------------------------------------------
# backward
# ----------------------------------------------------------------------------------------------
class _Grads:
    """fp32 gradient buffers of the parameters that require grad (zero-initialised; the kernels accumulate)."""

    def __init__(self):
        self.g: Dict[int, torch.Tensor] = {}

    def __call__(self, p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if p is None or not p.requires_grad:
            return None
        t = self.g.get(id(p))
        if t is None:
            t = self.g[id(p)] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
        return t


def _wt(mod, name: str, weight: torch.Tensor, fp16: bool, rows_pad=None) -> torch.Tensor:
    """Cached transposed 16-bit copy of a weight (B operand of its dgrad GEMM)."""
    cache = mod.__dict__.setdefault("_wt_cache", {})
    c = cache.get(name)
    if c is None:
        c = cache[name] = Weight16T(rows_pad)
    return c.get(weight)


def _linear_bwd(dy16, x16, lin, G, fp16, n_valid=None, k_valid=None, k_off=0, with_bias=True):
    """Parameter gradients of y = x W^T + b from dy (16-bit) and the saved operand x (16-bit).
    ``lin``: nn.Linear / nn.Conv1d(k=1) holding (weight, bias), or a LoraLinear (dense gradient projected
    onto the adapters: dA = s B^T dW, dB = s dW A^T -- parameter-space products of rank 16)."""
    pa = 1.0 / LOSS_SCALE[fp16]
    if isinstance(lin, LoraLinear):
        gA, gB = G(lin.A), G(lin.B)
        if gA is None and gB is None:
            return
        dw = torch.zeros(lin.out_features, lin.in_features, dtype=torch.float32, device=dy16.device)
        ops.wgrad(dy16, x16, dw, fp16, alpha=pa)
        if gA is not None:
            gA.addmm_(lin.B.detach().t(), dw, alpha=lin.scaling)
        if gB is not None:
            gB.addmm_(dw, lin.A.detach().t(), alpha=lin.scaling)
        return
    gb = G(lin.bias) if with_bias else None
    if gb is not None:
        ops.colsum16(dy16, gb, fp16, n_valid=n_valid, alpha=pa)
    gw = G(lin.weight)
    if gw is not None:
        ops.wgrad(dy16, x16, gw.view(gw.shape[0], -1), fp16, n_valid=n_valid, k_valid=k_valid, k_off=k_off, alpha=pa)


def _block_backward(blk, ws, roll, t: BlockTape, sc: Scratch, g32, G, fp16):
    """g32: gradient w.r.t. the block output, updated in place to the gradient w.r.t. the block input."""
    Z, H, W, C, T = ws.Z, ws.H, ws.W, ws.C, ws.T
    grid = (Z, H, W)
    mlp, att = blk.linear, blk.attention
    pa = 1.0 / LOSS_SCALE[fp16]
    ops.set_tag("hi" if C == 192 else "lo")
    # ---------------- x = x + s2 * LN2(Mlp(x))            (models/layers.py:251)
    # pre-LayerNorm values y2 = hidden W2^T + b2 (fp32 only: the plain-GEMM path of pangu_dgrad with the (out, in) weight)
    ops.dgrad(t.hidden, mlp._w2.get(mlp.linear2.weight), 0, fp16, out32=sc.y32, bias=mlp.linear2.bias)
    ops.layernorm_bwd(sc.y32, g32, blk.norm2.weight, G(blk.norm2.weight), G(blk.norm2.bias), T, C, LNB_IDENT, fp16,
                      dx16=sc.dy16, scale=t.s2, palpha=pa, dbias=G(mlp.linear2.bias))      # + d linear2.bias
    _linear_bwd(sc.dy16, t.hidden, mlp.linear2, G, fp16, with_bias=False)
    ops.dgrad(sc.dy16, _wt(mlp, "w2", mlp.linear2.weight, fp16), 1, fp16, out16=sc.dh16)
    ops.linear(t.xmid16, mlp._w1.get(mlp.linear1.weight), mlp.linear1.bias, None, sc.pre16, False, fp16)
    ops.gelu_bwd(sc.dh16, sc.pre16, fp16, dbias=G(mlp.linear1.bias), alpha=pa)                 # + d linear1.bias
    _linear_bwd(sc.dh16, t.xmid16, mlp.linear1, G, fp16, with_bias=False)
    ops.dgrad(sc.dh16, _wt(mlp, "w1", mlp.linear1.weight, fp16), 0, fp16, out32=g32, resid32=g32)
    # ---------------- x = shortcut + s1 * LN1(window_reverse(attention(window_partition(x))))   (:185-250)
    ops.dgrad(t.att, att._w2.get(att.linear2.weight), 0, fp16, out32=sc.y32, bias=att.linear2.bias)       # y1, fp32 only
    ops.layernorm_bwd(sc.y32, g32, blk.norm1.weight, G(blk.norm1.weight), G(blk.norm1.bias), T, C, LNB_IDENT, fp16,
                      dx16=sc.dy16, scale=t.s1, palpha=pa, dbias=G(att.linear2.bias))      # + d attention.linear2.bias
    _linear_bwd(sc.dy16, t.att, att.linear2, G, fp16, with_bias=False)
    dattw = sc.dattw[int(roll)]
    ops.dgrad(sc.dy16, _wt(att, "w2", att.linear2.weight, fp16), 3, fp16, out16=dattw, grid=grid, roll=roll)
    gbias = G(att.earth_specific_bias)       # None (frozen table, LoRA): the kernel skips the accumulation
    ops.window_attention_bwd(t.qkv, dattw, att.earth_specific_bias, sc.dqkv, gbias, Z, H, W, C, att.head_number,
                             roll, fp16, palpha=pa)
    # (the kernel can also emit d attention.linear1.bias -- dbqkv -- but its per-window shuffle reduction costs more
    #  than the separate column-sum pass over dqkv: +4.1 ms vs 2.0 ms per step, so the pass stays)
    _linear_bwd(sc.dqkv, t.xw, att.linear1, G, fp16)
    ops.dgrad(sc.dqkv, _wt(att, "w1", att.linear1.weight, fp16), 2, fp16, out32=g32, resid32=g32, grid=grid, roll=roll)
    ops.set_tag("")


def backward(model, tape: Tape, g_upper: torch.Tensor, g_surface: torch.Tensor, reducer=None) -> List[Optional[torch.Tensor]]:
    """Gradients of all parameters (in ``model.parameters()`` order; None where requires_grad is False)
    given dL/d(output), dL/d(output_surface).  ``reducer`` (``dist.GradReducer``): data-parallel gradient mean
    (era5_data/utils_dist.py:125-134), started group by group as the gradients become final."""
    hi, lo, fp16 = tape.hi, tape.lo, tape.fp16
    lat, lon = tape.lat, tape.lon
    plane = hi.H * hi.W
    G = _Grads()

    def done(module):          # every gradient of `module` is final: hand the group to the exchange
        if reducer is not None:
            reducer.ready([G.g.get(id(p)) for p in module.parameters()])
    sh, sl = tape.scr[id(hi)], tape.scr[id(lo)]
    f = torch.float32
    g_upper = g_upper.detach().to(f).contiguous()
    g_surface = g_surface.detach().to(f).contiguous()

    # ---------------- PatchRecovery + cat(skip, x)          (models/layers.py:511-545, pangu_model.py:81)
    rec = model._output_layer
    pa = 1.0 / LOSS_SCALE[fp16]
    ops.recover_grad_gather(g_upper, g_surface, tape.dyu, tape.dys, lat, lon, fp16, scale=LOSS_SCALE[fp16])
    for src, k_off in ((tape.skip16, 0), (tape.final16, 192)):
        _linear_bwd(tape.dyu, src[plane:], rec.conv, G, fp16, n_valid=160, k_off=k_off, with_bias=k_off == 0)
        _linear_bwd(tape.dys, src[:plane], rec.conv_surface, G, fp16, n_valid=64, k_off=k_off, with_bias=k_off == 0)
    wut = _wt(rec, "conv", rec.conv.weight, fp16, rows_pad=192)                   # [384, 192]
    wst = _wt(rec, "conv_surface", rec.conv_surface.weight, fp16, rows_pad=128)   # [384, 128]
    g_hi, g_skip = sh.g32, tape.g_skip
    ops.dgrad(tape.dyu, wut[:192], 0, fp16, out32=g_skip[plane:])
    ops.dgrad(tape.dys, wst[:192], 0, fp16, out32=g_skip[:plane])
    ops.dgrad(tape.dyu, wut[192:], 0, fp16, out32=g_hi[plane:])
    ops.dgrad(tape.dys, wst[192:], 0, fp16, out32=g_hi[:plane])
    done(rec)

    order, bt = tape.order, tape.blocks
    for b in (15, 14):
        blk, ws, roll = order[b]
        _block_backward(blk, ws, roll, bt[b], sh, g_hi, G, fp16)
        done(blk)

    # ---------------- UpSample                              (models/layers.py:474-499)
    up = model.upsample
    ops.cast_rows(g_hi, sh.dy16, fp16)
    _linear_bwd(sh.dy16, tape.up_a, up.linear2, G, fp16)
    ops.dgrad(sh.dy16, _wt(up, "w2", up.linear2.weight, fp16), 0, fp16, out32=sh.y32)
    ops.linear(tape.lo_out16, up._w1.get(up.linear1.weight), None, tape.u32, tape.u16, False, fp16)
    tape.du16.zero_()                                       # cropped positions (lat row 181) get no gradient
    ops.layernorm_bwd(tape.u32, sh.y32, up.norm.weight, G(up.norm.weight), G(up.norm.bias), hi.T, hi.C, LNB_UP, fp16,
                      dx16=tape.du16, grid=(hi.Z, hi.H, hi.W), palpha=pa)
    _linear_bwd(tape.du16, tape.lo_out16, up.linear1, G, fp16)
    g_lo = sl.g32
    ops.dgrad(tape.du16, _wt(up, "w1", up.linear1.weight, fp16), 0, fp16, out32=g_lo)
    done(up)

    for b in range(13, 1, -1):
        blk, ws, roll = order[b]
        _block_backward(blk, ws, roll, bt[b], sl, g_lo, G, fp16)
        done(blk)

    # ---------------- DownSample                            (models/layers.py:432-459)
    down = model.downsample
    ops.cast_rows(g_lo, sl.dy16, fp16)
    _linear_bwd(sl.dy16, tape.down_a, down.linear, G, fp16)
    ops.dgrad(sl.dy16, _wt(down, "w", down.linear.weight, fp16), 0, fp16, out32=tape.u32)
    ops.layernorm_bwd(tape.x32_skip, tape.u32, down.norm.weight, G(down.norm.weight), G(down.norm.bias), lo.T,
                      4 * hi.C, LNB_DOWN, fp16, dx32=g_skip, grid=(hi.Z, hi.H, hi.W), palpha=pa)
    done(down)

    for b in (1, 0):
        blk, ws, roll = order[b]
        _block_backward(blk, ws, roll, bt[b], sh, g_skip, G, fp16)
        done(blk)

    # ---------------- PatchEmbedding                        (models/layers.py:40-93; inputs need no gradient)
    emb = model._input_layer
    ops.cast_rows(g_skip, sh.dy16, fp16)
    _linear_bwd(sh.dy16[plane:], tape.a_u, emb.conv, G, fp16)
    _linear_bwd(sh.dy16[:plane], tape.a_s, emb.conv_surface, G, fp16, k_valid=112)
    done(emb)
    if reducer is not None:
        reducer.finish()
    return [G.g.get(id(p)) for p in model.parameters()]


class PanguTrainFunction(torch.autograd.Function):
    """Autograd hand-over: forward = ``forward_train``, backward = ``backward``.  The parameters are passed
    as inputs so that autograd routes the gradients into their ``.grad`` (accumulating, as usual)."""

    @staticmethod
    def forward(ctx, model, input, input_surface, statistics, maps, const_h, *params):
        for name, t in (("input", input), ("input_surface", input_surface)):
            if isinstance(t, torch.Tensor) and t.requires_grad:
                raise RuntimeError(f"pangu_pytorch_b200: gradients w.r.t. `{name}` are not produced by the B200 backward "
                                   "(the reference never asks for them either); detach it")
        with engine.operands(engine.training_operand_dtype()):
            out, out_s, tape = forward_train(model, input, input_surface, statistics, maps, const_h)
        ctx.model, ctx.tape, ctx.generation = model, tape, tape.generation
        return out, out_s

    @staticmethod
    def backward(ctx, g_upper, g_surface):
        if ctx.generation != ctx.tape.generation:
            raise RuntimeError("pangu_pytorch_b200: the saved activations of this forward were overwritten by a later "
                               "training-mode forward of the same model (one tape per model): call backward() before the "
                               "next forward, or run the extra forward under torch.no_grad() / model.eval()")
        with engine.operands("fp16" if ctx.tape.fp16 else "bf16"):
            grads = backward(ctx.model, ctx.tape, g_upper, g_surface, getattr(ctx.model, "grad_reducer", None))
        return (None,) * 6 + tuple(grads)


def train_step(model, input, input_surface, statistics, maps, const_h, target, target_surface):
    """Forward + weighted-L1 loss + backward of one sample, as the reference's training loop does
    (models/pangu_sample.py:52-69; the optimiser step is the caller's).  ``target*`` are PHYSICAL fields:
    ``normData`` is applied inside the loss kernel.  Gradients are accumulated into ``param.grad``.
    Returns the loss as a 1-element device tensor."""
    dev = model._input_layer.conv.weight.device
    out, out_s = model(input, input_surface, statistics, maps, const_h)
    s_mean, s_std = engine.f32(statistics[0], dev).reshape(4), engine.f32(statistics[1], dev).reshape(4)
    u_mean, u_std = engine.f32(statistics[2], dev).reshape(13, 5), engine.f32(statistics[3], dev).reshape(13, 5)
    loss, gu, gs = ops.l1_loss(out.detach(), out_s.detach(), engine.f32(target, dev), engine.f32(target_surface, dev),
                               s_mean, s_std, u_mean, u_std, want_grad=True)
    torch.autograd.backward((out, out_s), (gu, gs))
    return loss
