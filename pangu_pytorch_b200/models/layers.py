"""B200-native drop-in for the reference's ``models/layers.py``.

Same class names, constructor signatures, sub-module / parameter names (hence an identical
``state_dict``) and ``forward`` signatures as the reference; every forward is executed by
the hand-written sm_100a kernels behind ``include/pangu_b200.h``.  The ``nn.Linear`` /
``nn.Conv1d`` / ``nn.LayerNorm`` children only *hold* the fp32 parameters (so checkpoints,
optimisers and peft's ``isinstance(m, nn.Linear)`` discovery keep working); they are never
called on the hot path.

Inside ``PanguModel`` the blocks are chained through ``_run`` so that each kernel's epilogue
writes the 16-bit operand of the next kernel directly (already rolled and window
partitioned); the public ``forward`` methods accept / return plain fp32 tensors like the
reference and are what the module-level parity tests exercise.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

from .. import engine, ops
from ..engine import Weight16, workspace


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class DropPath(nn.Module):
    """Stochastic depth (timm semantics, as imported by the reference at models/layers.py:9).
    With batch 1 a draw is a single scalar: 0 or 1/keep.  ``scale()`` returns it so the factor
    can be folded into the LayerNorm+residual epilogue."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def scale(self) -> float:
        if self.drop_prob == 0.0 or not self.training:
            return 1.0
        keep = 1.0 - self.drop_prob
        draw = float(torch.empty(1).bernoulli_(keep).item())
        return draw / keep if (keep > 0.0 and self.scale_by_keep) else draw

    def forward(self, x):
        return x * self.scale()


class _Identity(nn.Identity):
    def scale(self) -> float:
        return 1.0


class PatchEmbedding_pretrain(nn.Module):
    """reference models/layers.py:12-93"""

    def __init__(self, patch_size, dim):
        super().__init__()
        self.conv = nn.Conv1d(in_channels=192, out_channels=dim, kernel_size=1, stride=1)
        self.conv_surface = nn.Conv1d(in_channels=112, out_channels=dim, kernel_size=1, stride=1)
        self.window_size = (2, 6, 12)
        self._w = Weight16()
        self._ws = Weight16(k_pad=128)

    def _run(self, input, input_surface, statistics, maps, const_h, ws: engine.GridWorkspace):
        dev = ws.x32.device
        fp16 = ws.fp16
        lat, lon = input_surface.shape[-2], input_surface.shape[-1]
        if input.shape[0] != 1:
            raise ValueError("batch size is 1 on this path, as in the reference (models/layers.py:219,227)")
        s_mean, s_std = engine.f32(statistics[0], dev).reshape(4), engine.f32(statistics[1], dev).reshape(4)
        u_mean, u_std = engine.f32(statistics[2], dev).reshape(13, 5), engine.f32(statistics[3], dev).reshape(13, 5)
        plane = ws.H * ws.W
        h = ops.dtype16(fp16)
        a_u = torch.empty(7 * plane, 192, dtype=h, device=dev)
        a_s = torch.empty(plane, 128, dtype=h, device=dev)
        ops.patch_embed(engine.f32(input, dev), engine.f32(input_surface, dev), s_mean, s_std, u_mean, u_std,
                        engine.f32(maps, dev), engine.f32(const_h, dev),
                        self._w.get(self.conv.weight), self.conv.bias, self._ws.get(self.conv_surface.weight),
                        self.conv_surface.bias, a_u, a_s, ws.x32, ws.x16w[0], lat, lon, fp16)

    def forward(self, input, input_surface, statistics, maps, const_h):
        lat, lon = input_surface.shape[-2], input_surface.shape[-1]
        ws = workspace(self.conv.weight.device, 8, (lat + 3) // 4, lon // 4, 192)
        self._run(input, input_surface, statistics, maps, const_h, ws)
        return ws.x32.clone().unsqueeze(0)


class Mlp(nn.Module):
    """reference models/layers.py:255-270"""

    def __init__(self, dim, dropout_rate):
        super().__init__()
        self.linear1 = nn.Linear(dim, dim * 4)
        self.linear2 = nn.Linear(dim * 4, dim)
        self.activation = nn.GELU()
        self.drop = nn.Dropout(dropout_rate)
        self._w1, self._w2 = Weight16(), Weight16()

    def forward(self, x):
        fp16 = engine.use_fp16()
        shape, C = x.shape, x.shape[-1]
        x2 = x.detach().reshape(-1, C).contiguous().float()
        M, h = x2.shape[0], ops.dtype16(fp16)
        a16 = torch.empty(M, C, dtype=h, device=x.device)
        _cast_rows(x2, a16, fp16)
        hid = torch.empty(M, 4 * C, dtype=h, device=x.device)
        ops.linear(a16, self._w1.get(self.linear1.weight), self.linear1.bias, None, hid, True, fp16)
        out32 = torch.empty(M, C, dtype=torch.float32, device=x.device)
        out16 = torch.empty(M, C, dtype=h, device=x.device)
        ops.linear(hid, self._w2.get(self.linear2.weight), self.linear2.bias, out32, out16, False, fp16)
        return out32.view(shape)


def _cast_rows(x32: torch.Tensor, out16: torch.Tensor, fp16: bool) -> None:
    """fp32 [M, C] -> 16-bit [M, C] with the cast kernel (stand-alone module API only)."""
    ops.cast_rows(x32, out16, fp16)


class EarthAttention3D(nn.Module):
    """reference models/layers.py:272-421"""

    def __init__(self, dim, heads, dropout_rate, window_size, device):
        super().__init__()
        self.device = device
        self.linear1 = nn.Linear(dim, dim * 3, bias=True)
        self.linear2 = nn.Linear(dim, dim)
        self.softmax = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout_rate)
        self.head_number = heads
        self.dim = dim
        self.scale = (dim // heads) ** -0.5
        self.window_size = window_size
        if dim == 192:
            input_shape = [8, 186]
        elif dim == 384:
            input_shape = [8, 96]
        else:
            raise ValueError("dim must be 192 or 384 (models/layers.py:298-301)")
        self.type_of_windows = (input_shape[0] // window_size[0]) * (input_shape[1] // window_size[1])
        n = window_size[0] * window_size[1] * window_size[2]
        self.earth_specific_bias = nn.Parameter(torch.zeros(1, self.type_of_windows, heads, n, n, device=device))
        _trunc_normal_(self.earth_specific_bias, std=0.02)
        self._construct_index()
        self._w1, self._w2 = Weight16(), Weight16()

    def _construct_index(self):
        """Compressed-bias index of the paper (models/layers.py:319-357); a plain attribute, unused
        by forward exactly as in the reference (the bias parameter is stored expanded)."""
        wz, wh, ww = self.window_size
        k = torch.arange(wz * wh * ww)
        z, h, w = k // (wh * ww), (k // ww) % wh, k % ww
        idx = (z[:, None] + z[None, :] * wz) * (2 * ww - 1) * wh * wh + (h[:, None] + h[None, :] * wh) * (2 * ww - 1) \
            + (w[:, None] - w[None, :] + ww - 1)
        self.position_index = idx.reshape(-1).to(self.device)

    def _run(self, ws: engine.GridWorkspace, roll: bool, window_order_out: bool = False):
        """x16w[roll] -> qkv (head-major, window-order rows) -> att (natural token order unless asked otherwise)."""
        fp16 = ws.fp16
        ops.qkv(ws.x16w[int(roll)], self._w1.get(self.linear1.weight), self.linear1.bias, ws.qkv,
                ws.Z, ws.H, ws.W, ws.C, fp16)
        ops.window_attention(ws.qkv, self.earth_specific_bias, ws.att, ws.Z, ws.H, ws.W, ws.C, self.head_number,
                             roll, fp16, window_order_out)

    def forward(self, x, mask):
        """x: [nLon, types, 144, C] window tensor.  ``mask`` is only inspected for None-ness: the
        kernel regenerates the reference's shifted-window mask (gen_mask) in registers."""
        nlon, types, n, C = x.shape
        H = {192: 181, 384: 91}[C]
        ws = workspace(x.device, 8, H, 12 * nlon, C)
        fp16 = ws.fp16
        x2 = x.detach().reshape(-1, C).contiguous().float()
        roll = mask is not None
        _cast_rows(x2, ws.x16w[int(roll)], fp16)
        self._run(ws, roll, window_order_out=True)
        out32 = torch.empty(ws.Tp, C, dtype=torch.float32, device=x.device)
        out16 = torch.empty(ws.Tp, C, dtype=ops.dtype16(fp16), device=x.device)
        ops.linear(ws.att, self._w2.get(self.linear2.weight), self.linear2.bias, out32, out16, False, fp16)
        # this stand-alone path overwrote the static zero pad rows of x16w: restore them
        ws.x16w[int(roll)].zero_()
        return out32.view(x.shape)


class EarthSpecificBlock(nn.Module):
    """reference models/layers.py:127-253"""

    def __init__(self, dim, drop_path_ratio, heads, device):
        super().__init__()
        self.device = device
        self.window_size = (2, 6, 12)
        self.drop_path = DropPath(drop_path_ratio) if drop_path_ratio > 0.0 else _Identity()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.linear = Mlp(dim, 0)
        self.attention = EarthAttention3D(dim, heads, 0, self.window_size, device=self.device)
        self.padding_front, self.padding_back = 0, 5
        self.type_of_windows = self.attention.type_of_windows
        self.dim = dim

    def gen_mask(self, x):
        """Shifted-window mask (models/layers.py:153-181) in closed form; returned
        for API parity -- the attention kernel computes the same predicate in registers.
        x: rolled, padded [1, Z, Hp, W, C]."""
        _, Z, Hp, W, _ = x.shape
        nZ, nH, nLon = Z // 2, Hp // 6, W // 12
        k = torch.arange(144, device=x.device)
        zl, hl = k // 72, (k // 12) % 6
        m = torch.zeros(nZ, nH, 144, 144, dtype=torch.bool, device=x.device)
        m[nZ - 1] |= zl[:, None] != zl[None, :]
        m[:, nH - 1] |= (hl[:, None] < 3) != (hl[None, :] < 3)
        mask = torch.zeros(nZ * nH, 144, 144, device=x.device)
        mask[m.view(nZ * nH, 144, 144)] = -100.0
        return mask.unsqueeze(0).expand(nLon, -1, -1, -1)

    def _run(self, ws: engine.GridWorkspace, roll: bool, roll_out: int, x16_out=None):
        """One block on the workspace.  Expects ws.x32 and ws.x16w[roll]; leaves the new ws.x32 and
        its 16-bit shadow either in ws.x16w[roll_out] (window order for the next block) or, for
        roll_out < 0, in ``x16_out`` (natural order; default ws.x16)."""
        fp16 = ws.fp16
        Z, H, W, C = ws.Z, ws.H, ws.W, ws.C
        ops.set_tag("hi" if C == 192 else "lo")
        s1, s2 = self.drop_path.scale(), self.drop_path.scale()
        att = self.attention
        att._run(ws, roll)
        ops.proj_ln_residual(ws.att, att._w2.get(att.linear2.weight), att.linear2.bias, self.norm1.weight,
                             self.norm1.bias, ws.x32, ws.x16, Z, H, W, C, roll, s1, fp16)
        mlp = self.linear
        target = ws.x16w[roll_out] if roll_out >= 0 else (ws.x16 if x16_out is None else x16_out)
        ops.mlp_ln_residual(ws.x16, mlp._w1.get(mlp.linear1.weight), mlp.linear1.bias,
                            mlp._w2.get(mlp.linear2.weight), mlp.linear2.bias, self.norm2.weight, self.norm2.bias,
                            ws.hidden, ws.x32, target, Z, H, W, C, roll_out, s2, fp16)
        ops.set_tag("")
        return s1, s2

    def forward(self, x, Z, H, W, roll):
        C = x.shape[-1]
        ws = workspace(x.device, Z, H, W, C)
        ws.x32.copy_(x.detach().reshape(ws.T, C))
        ops.to_window16(ws.x32, ws.x16w[int(roll)], Z, H, W, C, int(roll), ws.fp16)
        self._run(ws, bool(roll), -1)
        return ws.x32.clone().view(x.shape)


class EarthSpecificLayer(nn.Module):
    """reference models/layers.py:96-125.  ``use_checkpoint`` is accepted and stored for API
    parity; activation checkpointing is an autograd-memory device of the reference and has no
    effect on the values computed."""

    def __init__(self, depth, dim, drop_path_ratio_list, heads, use_checkpoint, device):
        super().__init__()
        self.device = device
        self.depth = depth
        block_list = OrderedDict()
        for i in range(depth):
            block_list["EarthSpecificBlock{}".format(i)] = EarthSpecificBlock(dim, drop_path_ratio_list[i], heads,
                                                                              device=device)
        self.blocks = nn.Sequential(block_list)
        self.use_checkpoint = use_checkpoint

    def _run(self, ws: engine.GridWorkspace, last_roll_out: int, x16_out=None):
        """Blocks alternate roll = i % 2 (models/layers.py:116-124); the last block hands its 16-bit
        shadow over as requested by the caller."""
        n = len(self.blocks)
        for i, blk in enumerate(self.blocks):
            if i + 1 < n:
                blk._run(ws, i % 2 == 1, (i + 1) % 2)
            else:
                blk._run(ws, i % 2 == 1, last_roll_out, x16_out)

    def forward(self, x, Z, H, W):
        C = x.shape[-1]
        ws = workspace(x.device, Z, H, W, C)
        ws.x32.copy_(x.detach().reshape(ws.T, C))
        ops.to_window16(ws.x32, ws.x16w[0], Z, H, W, C, 0, ws.fp16)
        self._run(ws, -1)
        return ws.x32.clone().view(x.shape)


class DownSample(nn.Module):
    """reference models/layers.py:423-459"""

    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(in_features=4 * dim, out_features=2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)
        self._w = Weight16()

    def _run(self, hi: engine.GridWorkspace, lo: engine.GridWorkspace):
        h = ops.dtype16(hi.fp16)
        ws_a = torch.empty(lo.T, 4 * hi.C, dtype=h, device=hi.x32.device)
        ops.downsample(hi.x32, self.norm.weight, self.norm.bias, self._w.get(self.linear.weight), ws_a, lo.x32,
                       lo.x16w[0], hi.Z, hi.H, hi.W, hi.C, hi.fp16)

    def forward(self, x, Z, H, W):
        C = x.shape[-1]
        hi = workspace(x.device, Z, H, W, C)
        lo = workspace(x.device, Z, (H + 1) // 2, W // 2, 2 * C)
        hi.x32.copy_(x.detach().reshape(hi.T, C))
        self._run(hi, lo)
        return lo.x32.clone().unsqueeze(0)


class UpSample(nn.Module):
    """reference models/layers.py:461-499 (the reference hard-codes the (8, 91, 180) input grid;
    here the longitude extent is inferred from the token count)."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.linear1 = nn.Linear(input_dim, output_dim * 4, bias=False)
        self.linear2 = nn.Linear(output_dim, output_dim, bias=False)
        self.norm = nn.LayerNorm(output_dim)
        self._w1, self._w2 = Weight16(), Weight16()

    def _run(self, lo: engine.GridWorkspace, hi: engine.GridWorkspace):
        h = ops.dtype16(hi.fp16)
        ws_a = torch.empty(hi.T, hi.C, dtype=h, device=hi.x32.device)
        ops.upsample(lo.x16, self._w1.get(self.linear1.weight), self.norm.weight, self.norm.bias,
                     self._w2.get(self.linear2.weight), ws_a, hi.x32, hi.x16w[0], hi.Z, hi.H, hi.W, lo.C, hi.C,
                     hi.fp16)

    def forward(self, x, Z=8, H=181):
        C2 = x.shape[-1]
        H2 = (H + 1) // 2
        W2 = x.shape[1] // (Z * H2)
        lo = workspace(x.device, Z, H2, W2, C2)
        hi = workspace(x.device, Z, H, 2 * W2, C2 // 2)
        lo.x32.copy_(x.detach().reshape(lo.T, C2))
        _cast_rows(lo.x32, lo.x16, lo.fp16)
        self._run(lo, hi)
        return hi.x32.clone().unsqueeze(0)


class PatchRecovery_pretrain(nn.Module):
    """reference models/layers.py:501-545"""

    def __init__(self, dim):
        super().__init__()
        self.patch_size = (2, 4, 4)
        self.dim = dim
        self.conv = nn.Conv1d(in_channels=dim, out_channels=160, kernel_size=1, stride=1)
        self.conv_surface = nn.Conv1d(in_channels=dim, out_channels=64, kernel_size=1, stride=1)
        self._w, self._ws = Weight16(), Weight16()

    def _run(self, skip16, hi: engine.GridWorkspace, lat: int, lon: int):
        dev = hi.x32.device
        out = torch.empty(1, 5, 13, lat, lon, dtype=torch.float32, device=dev)
        out_s = torch.empty(1, 4, lat, lon, dtype=torch.float32, device=dev)
        ops.patch_recover(skip16, hi.x16, self._w.get(self.conv.weight), self.conv.bias,
                          self._ws.get(self.conv_surface.weight), self.conv_surface.bias, out, out_s,
                          hi.Z, hi.H, hi.W, hi.C, lat, lon, hi.fp16)
        return out, out_s

    def forward(self, x, Z, H, W, lat=None):
        C = x.shape[-1] // 2
        hi = workspace(x.device, Z, H, W, C)
        lat = (4 * H - 3) if lat is None else lat          # 721 for H = 181 (models/layers.py:527-529)
        x2 = x.detach().reshape(hi.T, 2 * C).float()
        skip16 = torch.empty(hi.T, C, dtype=ops.dtype16(hi.fp16), device=x.device)
        _cast_rows(x2[:, :C].contiguous(), skip16, hi.fp16)
        _cast_rows(x2[:, C:].contiguous(), hi.x16, hi.fp16)
        return self._run(skip16, hi, lat, 4 * W)
