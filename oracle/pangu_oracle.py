"""CPU fp32 restatement of the Pangu-Weather hot path (TEST INFRASTRUCTURE ONLY).

This file is the parity oracle for the B200 kernels.  It is *not* part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``pangu_pytorch_b200``) never imports anything under ``oracle/``.

What it restates (all citations relative to the reference checkout):

* ``models/layers.py:40-93``    PatchEmbedding_pretrain.forward   -> ``patch_embed``
* ``models/layers.py:153-181``  EarthSpecificBlock.gen_mask       -> ``shift_mask``
* ``models/layers.py:183-253``  EarthSpecificBlock.forward        -> ``earth_block``
* ``models/layers.py:264-270``  Mlp.forward                       -> inside ``earth_block``
* ``models/layers.py:360-421``  EarthAttention3D.forward          -> ``window_attention``
* ``models/layers.py:432-459``  DownSample.forward                -> ``down_sample``
* ``models/layers.py:474-499``  UpSample.forward                  -> ``up_sample``
* ``models/layers.py:511-545``  PatchRecovery_pretrain.forward    -> ``patch_recover``
* ``models/pangu_model.py:50-87`` PanguModel.forward              -> ``forward``
* ``models/pangu_sample.py:57-67`` weighted L1 loss               -> ``weighted_l1_loss``
* ``era5_data/utils_data.py:315-330`` normData / normBackData     -> ``norm_data`` / ``norm_back_data``
* ``era5_data/utils_dist.py:125-134`` gather_grad                 -> ``mean_of_grads``
* ``finetune/lora_tune.py:124-139`` peft LoRA (third party, unpinned) -> ``lora_linear``
* ``era5_data/score.py:83-135`` latitude-weighted RMSE / ACC    -> ``weighted_rmse_channels`` / ``weighted_acc_channels``
* ``models/pangu_sample.py:203-270`` evaluation loop scores       -> ``evaluation_scores``

Unlike the reference (which chains view/permute/roll/pad and hard-codes 181x360 in
three places) the oracle is written against *closed-form index maps* (SURVEY.md
Appendix A) and is generic in the longitude extent, so that a full-depth forward can
be evaluated on a narrow longitude strip in seconds.  Parity is pinned: the reference
itself is imported and executed in the build container by ``oracle/make_golden.py``
and the sampled outputs are committed under ``tests/golden``; ``tests/test_oracle.py``
checks this file against those fixtures (and, when ``/root/reference`` is present,
against the live reference modules).

Weights are passed as a plain ``dict`` with the reference's ``state_dict`` keys.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

WINDOW = (2, 6, 12)          # models/layers.py:137
WIN_TOKENS = 2 * 6 * 12      # 144
SHIFT = (1, 3, 6)            # models/layers.py:201 (half windows)
PAD_LAT = 5                  # models/layers.py:145 padding_back, both resolutions
MASK_VALUE = -100.0          # models/layers.py:179
HEAD_DIM = 32
UPPER_WEIGHTS = (3.00, 0.60, 1.50, 0.77, 0.54)   # era5_data/config.py:45
SURFACE_WEIGHTS = (1.50, 0.77, 0.66, 3.00)       # era5_data/config.py:46
DEPTHS = (2, 6, 6, 2)        # models/pangu_model.py:9
HEADS = (6, 12, 12, 6)
DIMS = (192, 384, 384, 192)


# --------------------------------------------------------------------------------------
# integer contracts (bit-exact)
# --------------------------------------------------------------------------------------
def window_source_index(Z: int, H: int, W: int, roll: bool) -> Tensor:
    """Window-ordered gather map of ``EarthSpecificBlock`` (models/layers.py:188-221).

    Returns int64 ``[nLon, types, 144]``; entry = flat token index ``(z*H + h)*W + w`` of
    the un-padded ``(Z, H, W)`` grid, or ``-1`` for a zero pad token (lat rows H..H+4,
    which ride along with the roll).  SURVEY.md Appendix A1.
    """
    Hp = H + PAD_LAT
    assert Z % 2 == 0 and Hp % 6 == 0 and W % 12 == 0
    nZ, nH, nLon = Z // 2, Hp // 6, W // 12
    s = SHIFT if roll else (0, 0, 0)
    lw = torch.arange(nLon).view(nLon, 1, 1, 1, 1, 1)
    zw = torch.arange(nZ).view(1, nZ, 1, 1, 1, 1)
    hw = torch.arange(nH).view(1, 1, nH, 1, 1, 1)
    zl = torch.arange(2).view(1, 1, 1, 2, 1, 1)
    hl = torch.arange(6).view(1, 1, 1, 1, 6, 1)
    wl = torch.arange(12).view(1, 1, 1, 1, 1, 12)
    zp = (2 * zw + zl + s[0]) % Z
    hp = (6 * hw + hl + s[1]) % Hp
    wp = (12 * lw + wl + s[2]) % W
    src = (zp * H + hp) * W + wp
    src = torch.where(hp >= H, torch.full_like(src, -1), src)
    return src.reshape(nLon, nZ * nH, WIN_TOKENS)


def shift_mask(Z: int, H: int) -> Tensor:
    """Shifted-window mask of ``gen_mask`` (models/layers.py:153-181), ``[types,144,144]``
    of {0,-100}; identical for every longitude window (W is cyclic, never masked).
    SURVEY.md Appendix A2."""
    Hp = H + PAD_LAT
    nZ, nH = Z // 2, Hp // 6
    k = torch.arange(WIN_TOKENS)
    zl, hl = k // 72, (k // 12) % 6
    z_split = (zl[:, None] != zl[None, :])
    h_split = ((hl[:, None] < 3) != (hl[None, :] < 3))
    m = torch.zeros(nZ, nH, WIN_TOKENS, WIN_TOKENS, dtype=torch.bool)
    m[nZ - 1] |= z_split
    m[:, nH - 1] |= h_split
    out = torch.zeros(nZ * nH, WIN_TOKENS, WIN_TOKENS)
    out[m.view(nZ * nH, WIN_TOKENS, WIN_TOKENS)] = MASK_VALUE
    return out


def position_index() -> Tensor:
    """``EarthAttention3D._construct_index`` (models/layers.py:319-357): the (unused on
    the forward path) compressed-bias index, int64 [20736] in 0..3311."""
    wz, wh, ww = WINDOW
    k = torch.arange(WIN_TOKENS)
    z, h, w = k // (wh * ww), (k // ww) % wh, k % ww
    dz = z[:, None] + z[None, :] * wz
    dh = h[:, None] + h[None, :] * wh
    dw = w[:, None] - w[None, :] + ww - 1
    return (dz * (2 * ww - 1) * wh * wh + dh * (2 * ww - 1) + dw).reshape(-1)


# --------------------------------------------------------------------------------------
# floating-point modules
# --------------------------------------------------------------------------------------
def window_attention(xw: Tensor, p: Dict[str, Tensor], pre: str, heads: int,
                     mask: Optional[Tensor]) -> Tensor:
    """``EarthAttention3D.forward`` (models/layers.py:360-421).  ``xw``: [nLon,types,144,C]."""
    nLon, types, N, C = xw.shape
    qkv = F.linear(xw, p[pre + "linear1.weight"], p[pre + "linear1.bias"])
    qkv = qkv.view(nLon, types, N, 3, heads, C // heads).permute(3, 0, 1, 4, 2, 5)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    s = q @ k.transpose(-2, -1)                        # [nLon,types,heads,144,144]
    s = s + p[pre + "earth_specific_bias"]             # [1,types,heads,144,144]
    if mask is not None:
        s = s + mask.view(1, types, 1, N, N)
    a = torch.softmax(s, dim=-1)
    o = (a @ v).permute(0, 1, 3, 2, 4).reshape(nLon, types, N, C)
    return F.linear(o, p[pre + "linear2.weight"], p[pre + "linear2.bias"])


def earth_block(x: Tensor, p: Dict[str, Tensor], pre: str, Z: int, H: int, W: int,
                heads: int, roll: bool, drop_scale=1.0) -> Tensor:
    """``EarthSpecificBlock.forward`` (models/layers.py:183-253), eval-mode DropPath.

    ``x``: [1, Z*H*W, C].  ``drop_scale`` multiplies both residual branches (a DropPath
    draw for batch 1 is a single Bernoulli/keep scalar, so training-mode parity can be
    exercised by passing 0 or 1/keep; a pair ``(s1, s2)`` gives the two DropPath draws of
    models/layers.py:250-251 separately)."""
    s1, s2 = drop_scale if isinstance(drop_scale, (tuple, list)) else (drop_scale, drop_scale)
    C = x.shape[-1]
    src = window_source_index(Z, H, W, roll)           # [nLon,types,144]
    flat = src.reshape(-1)
    xt = x.reshape(-1, C)
    real = flat >= 0
    xw = torch.zeros(flat.numel(), C, dtype=x.dtype).index_put((real.nonzero().squeeze(1),), xt[flat[real]])
    xw = xw.view(*src.shape, C)
    mask = shift_mask(Z, H) if roll else None
    aw = window_attention(xw, p, pre + "attention.", heads, mask).reshape(-1, C)
    # window reverse + un-roll + crop == scatter through the same map; pad rows dropped
    y = torch.zeros_like(xt).index_put((flat[real],), aw[real])
    y = F.layer_norm(y, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
    xt = xt + s1 * y
    h = F.linear(xt, p[pre + "linear.linear1.weight"], p[pre + "linear.linear1.bias"])
    h = F.gelu(h)                                       # exact erf GELU (nn.GELU default)
    h = F.linear(h, p[pre + "linear.linear2.weight"], p[pre + "linear.linear2.bias"])
    h = F.layer_norm(h, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    return (xt + s2 * h).view(1, -1, C)


def earth_layer(x, p, layer: int, Z, H, W, depth: int, heads: int, drop_scales=None) -> Tensor:
    """``EarthSpecificLayer.forward`` (models/layers.py:110-125): roll every odd block.
    ``drop_scales``: optional per-block DropPath factors (train mode), else 1 (eval)."""
    for i in range(depth):
        pre = f"layers.EarthSpecificLayer{layer}.blocks.EarthSpecificBlock{i}."
        x = earth_block(x, p, pre, Z, H, W, heads, roll=(i % 2 == 1),
                        drop_scale=1.0 if drop_scales is None else drop_scales[i])
    return x


def embed_operands(upper: Tensor, surface: Tensor, statistics: Sequence[Tensor],
                   maps: Tensor, const_h: Tensor) -> Tuple[Tensor, Tensor]:
    """im2col operands of ``PatchEmbedding_pretrain`` (models/layers.py:48-85), SURVEY A4.

    Returns ``A_surface [Hh*Ww, 112]`` (feature ``(c*4+dh)*4+dw``) and
    ``A_upper [7*Hh*Ww, 192]`` (feature ``((c*2+dz)*4+dh)*4+dw``)."""
    s_mean, s_std, u_mean, u_std = statistics
    lat, lon = surface.shape[-2], surface.shape[-1]
    Hh, Ww = (lat + 3) // 4, lon // 4
    sf = (surface[0].permute(1, 2, 0) - s_mean.reshape(4)) / s_std.reshape(4)   # [lat,lon,4]
    sf = F.pad(sf.permute(2, 0, 1), (0, 0, 0, 4 * Hh - lat))                    # [4,724,lon]
    sf = torch.cat((sf, maps[0]), dim=0)                                        # [7,724,lon]
    a_s = sf.view(7, Hh, 4, Ww, 4).permute(1, 3, 0, 2, 4).reshape(Hh * Ww, 112)
    # level l is normalised with upper_mean[12-l] (the two flips at layers.py:73,76)
    um = u_mean.reshape(13, 5).flip(0).t().reshape(5, 13, 1, 1)
    us = u_std.reshape(13, 5).flip(0).t().reshape(5, 13, 1, 1)
    up = (upper[0] - um) / us                                                   # [5,13,lat,lon]
    up = torch.cat((up, const_h.reshape(1, 13, lat, lon)), dim=0)               # [6,13,lat,lon]
    up = F.pad(up, (0, 0, 0, 4 * Hh - lat, 0, 1))                               # [6,14,724,lon]
    a_u = up.view(6, 7, 2, Hh, 4, Ww, 4).permute(1, 3, 5, 0, 2, 4, 6).reshape(7 * Hh * Ww, 192)
    return a_s, a_u


def patch_embed(upper, surface, statistics, maps, const_h, p) -> Tensor:
    """``PatchEmbedding_pretrain.forward`` (models/layers.py:40-93) -> [1, 8*Hh*Ww, 192]."""
    a_s, a_u = embed_operands(upper, surface, statistics, maps, const_h)
    xs = F.linear(a_s, p["_input_layer.conv_surface.weight"][:, :, 0], p["_input_layer.conv_surface.bias"])
    xu = F.linear(a_u, p["_input_layer.conv.weight"][:, :, 0], p["_input_layer.conv.bias"])
    return torch.cat((xs, xu), dim=0).unsqueeze(0)


def down_sample(x: Tensor, p, Z: int, H: int, W: int) -> Tensor:
    """``DownSample.forward`` (models/layers.py:432-459), SURVEY A6."""
    C = x.shape[-1]
    x = F.pad(x.view(Z, H, W, C), (0, 0, 0, 0, 0, H % 2))
    H2, W2 = (H + 1) // 2, W // 2
    x = x.view(Z, H2, 2, W2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(Z * H2 * W2, 4 * C)
    x = F.layer_norm(x, (4 * C,), p["downsample.norm.weight"], p["downsample.norm.bias"], 1e-5)
    return F.linear(x, p["downsample.linear.weight"]).unsqueeze(0)


def up_sample(x: Tensor, p, Z: int, H2: int, W2: int, H: int) -> Tensor:
    """``UpSample.forward`` (models/layers.py:474-499), SURVEY A6."""
    x = F.linear(x.reshape(Z * H2 * W2, -1), p["upsample.linear1.weight"])
    C = x.shape[-1] // 4
    x = x.view(Z, H2, W2, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(Z, 2 * H2, 2 * W2, C)
    x = x[:, :H].reshape(Z * H * 2 * W2, C)
    x = F.layer_norm(x, (C,), p["upsample.norm.weight"], p["upsample.norm.bias"], 1e-5)
    return F.linear(x, p["upsample.linear2.weight"]).unsqueeze(0)


def patch_recover(x: Tensor, p, Z: int, H: int, W: int, lat: int) -> Tuple[Tensor, Tensor]:
    """``PatchRecovery_pretrain.forward`` (models/layers.py:511-545), SURVEY A5.
    ``x``: [1, Z*H*W, 384]; returns normalised ``(1,5,13,lat,4W)``, ``(1,4,lat,4W)``."""
    xt = x.reshape(Z, H * W, -1)
    yu = F.linear(xt[1:].reshape((Z - 1) * H * W, -1), p["_output_layer.conv.weight"][:, :, 0],
                  p["_output_layer.conv.bias"])
    yu = yu.view(Z - 1, H, W, 5, 2, 4, 4).permute(3, 0, 4, 1, 5, 2, 6).reshape(5, 2 * (Z - 1), 4 * H, 4 * W)
    ys = F.linear(xt[0], p["_output_layer.conv_surface.weight"][:, :, 0], p["_output_layer.conv_surface.bias"])
    ys = ys.view(H, W, 4, 4, 4).permute(2, 0, 3, 1, 4).reshape(4, 4 * H, 4 * W)
    return yu[:, :13, :lat].unsqueeze(0).contiguous(), ys[:, :lat].unsqueeze(0).contiguous()


def forward(p: Dict[str, Tensor], upper, surface, statistics, maps, const_h,
            taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """``PanguModel.forward`` (models/pangu_model.py:50-87).  Generic in longitude: the
    number of longitude tokens ``lon/4`` must be a multiple of 24.  ``taps`` (optional
    dict) receives the residual stream after every stage."""
    with torch.no_grad():
        return forward_train(p, upper, surface, statistics, maps, const_h, taps=taps)


def forward_train(p: Dict[str, Tensor], upper, surface, statistics, maps, const_h,
                  drop_scales=None, taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """Differentiable ``PanguModel.forward`` (autograd records it when ``p`` holds leaves that
    require grad): the training path of models/pangu_sample.py:52.  ``drop_scales``: 16 pairs
    ``(s1, s2)`` of DropPath factors in block order (None = eval mode)."""
    lat, lon = surface.shape[-2], surface.shape[-1]
    Z, H, W = 8, (lat + 3) // 4, lon // 4
    H2, W2 = (H + 1) // 2, W // 2

    def tap(name, t):
        if taps is not None:
            taps[name] = t

    ds = [None] * 4 if drop_scales is None else [drop_scales[0:2], drop_scales[2:8], drop_scales[8:14], drop_scales[14:16]]
    x = patch_embed(upper, surface, statistics, maps, const_h, p); tap("embed", x)
    x = earth_layer(x, p, 0, Z, H, W, DEPTHS[0], HEADS[0], ds[0]); tap("layer0", x)
    skip = x
    x = down_sample(x, p, Z, H, W); tap("down", x)
    x = earth_layer(x, p, 1, Z, H2, W2, DEPTHS[1], HEADS[1], ds[1]); tap("layer1", x)
    x = earth_layer(x, p, 2, Z, H2, W2, DEPTHS[2], HEADS[2], ds[2]); tap("layer2", x)
    x = up_sample(x, p, Z, H2, W2, H); tap("up", x)
    x = earth_layer(x, p, 3, Z, H, W, DEPTHS[3], HEADS[3], ds[3]); tap("layer3", x)
    x = torch.cat((skip, x), dim=-1)
    return patch_recover(x, p, Z, H, W, lat)


def loss_and_grads(p: Dict[str, Tensor], upper, surface, statistics, maps, const_h, tgt_upper, tgt_surface,
                   drop_scales=None):
    """One training step's loss and parameter gradients (models/pangu_sample.py:52-69): forward,
    ``normData`` of the physical targets, weighted L1, ``loss.backward()``.
    Returns (loss, {name: grad}, (dL/d output, dL/d output_surface))."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    ou, os_ = forward_train(leaves, upper, surface, statistics, maps, const_h, drop_scales)
    ou.retain_grad(); os_.retain_grad()
    tu, ts = norm_data(tgt_upper, tgt_surface, output_statistics(statistics))
    loss = weighted_l1_loss(ou, os_, tu, ts)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaves.items()}, (ou.grad, os_.grad)


# --------------------------------------------------------------------------------------
# glue on either side of the model
# --------------------------------------------------------------------------------------
def norm_data(upper, surface, statistics):
    """``normData`` (era5_data/utils_data.py:315-321); output-order statistics
    ``(1,4,1,1)``/``(1,5,13,1,1)`` as built by ``weatherStatistics_output`` (:214-236)."""
    s_mean, s_std, u_mean, u_std = statistics
    return (upper - u_mean) / u_std, (surface - s_mean) / s_std


def norm_back_data(upper, surface, statistics):
    """``normBackData`` (era5_data/utils_data.py:324-330)."""
    s_mean, s_std, u_mean, u_std = statistics
    return upper * u_std + u_mean, surface * s_std + s_mean


def output_statistics(statistics):
    """Input-order stats ``(4,),(4,),(13,1,1,5),(13,1,1,5)`` -> output-order stats
    (level-reversed, ``weatherStatistics_output`` era5_data/utils_data.py:214-236)."""
    s_mean, s_std, u_mean, u_std = statistics
    um = u_mean.reshape(13, 5).flip(0).t().reshape(1, 5, 13, 1, 1)
    us = u_std.reshape(13, 5).flip(0).t().reshape(1, 5, 13, 1, 1)
    return s_mean.reshape(1, 4, 1, 1), s_std.reshape(1, 4, 1, 1), um, us


def weighted_l1_loss(out_u, out_s, tgt_u, tgt_s) -> Tensor:
    """Training loss (models/pangu_sample.py:61-67, weights era5_data/config.py:45-46);
    targets already normalised by ``norm_data``."""
    wu = torch.tensor(UPPER_WEIGHTS, dtype=out_u.dtype).view(1, 5, 1, 1, 1)
    ws = torch.tensor(SURFACE_WEIGHTS, dtype=out_s.dtype).view(1, 4, 1, 1)
    return ((out_u - tgt_u).abs() * wu).mean() + 0.25 * ((out_s - tgt_s).abs() * ws).mean()


def latitude_weights(num_lat: int) -> Tensor:
    """``latitude_weighting_factor_torch`` (era5_data/score.py:83-90; the reference writes 3.1416 for pi)."""
    j = torch.arange(num_lat, dtype=torch.float64)
    c = torch.cos(3.1416 / 180.0 * (90.0 - j * 180.0 / float(num_lat - 1)))
    return num_lat * c / c.sum()


def weighted_rmse_channels(pred: Tensor, target: Tensor) -> Tensor:
    """``weighted_rmse_torch_channels`` (era5_data/score.py:92-105): ``[..., lat, lon]`` -> ``[...]``."""
    w = latitude_weights(pred.shape[-2]).to(pred.dtype).view(-1, 1)
    return torch.sqrt(torch.mean(w * (pred - target) ** 2.0, dim=(-1, -2)))


def weighted_acc_channels(pred: Tensor, target: Tensor) -> Tensor:
    """``weighted_acc_torch_channels`` (era5_data/score.py:123-135) on anomalies."""
    w = latitude_weights(pred.shape[-2]).to(pred.dtype).view(-1, 1)
    return torch.sum(w * pred * target, dim=(-1, -2)) / torch.sqrt(
        torch.sum(w * pred * pred, dim=(-1, -2)) * torch.sum(w * target * target, dim=(-1, -2)))


def evaluation_scores(out_u, out_s, tgt_u, tgt_s, statistics):
    """Scores of the reference's test loop (models/pangu_sample.py:203-270): ``normBackData`` of the model outputs,
    RMSE per (variable, level), ACC on anomalies against the scalar statistics means.  Returns
    (rmse_upper [5,13], rmse_surface [4], acc_upper [5,13], acc_surface [4])."""
    ostats = output_statistics(statistics)
    pu, ps = norm_back_data(out_u, out_s, ostats)
    s_mean, _, u_mean, _ = ostats
    ru, rs = weighted_rmse_channels(pu[0], tgt_u[0]), weighted_rmse_channels(ps[0], tgt_s[0])
    au = weighted_acc_channels(pu[0] - u_mean[0], tgt_u[0] - u_mean[0])
    as_ = weighted_acc_channels(ps[0] - s_mean[0], tgt_s[0] - s_mean[0])
    return ru, rs, au, as_


def rollout(p, upper, surface, statistics, maps, const_h, steps: int):
    """Autoregressive chain (inference/inference_singleOutput.py:92-105 semantics; the torch
    model emits normalised fields so ``normBackData`` sits between steps)."""
    outs = []
    ostats = output_statistics(statistics)
    for _ in range(steps):
        ou, os_ = forward(p, upper, surface, statistics, maps, const_h)
        upper, surface = norm_back_data(ou, os_, ostats)
        outs.append((upper, surface))
    return outs


def mean_of_grads(per_rank_grads: Sequence[Sequence[Tensor]]):
    """``gather_grad`` (era5_data/utils_dist.py:125-134): all_reduce(SUM) then / world."""
    world = len(per_rank_grads)
    return [sum(g[i] for g in per_rank_grads) / world for i in range(len(per_rank_grads[0]))]


def lora_linear(x, weight, bias, lora_A, lora_B, alpha: float = 16.0, r: int = 16):
    """peft LoRA Linear as configured at finetune/lora_tune.py:129-139 (peft is third
    party and unpinned; dropout omitted = eval mode): ``W x + b + (alpha/r) B(A(x))``."""
    return F.linear(x, weight, bias) + (alpha / r) * F.linear(F.linear(x, lora_A), lora_B)


# --------------------------------------------------------------------------------------
# seeded synthetic weights / inputs -- pure torch, no reference import
# --------------------------------------------------------------------------------------
def param_shapes() -> "list[tuple[str, tuple]]":
    """The 223 ``state_dict`` entries in the reference's registration order
    (models/pangu_model.py:16-38, models/layers.py:17-18,141-144,259-260,281-282,311,
    428-429,466-472,508-509; cross-checked against keys_all.csv)."""
    out = [("_input_layer.conv.weight", (192, 192, 1)), ("_input_layer.conv.bias", (192,)),
           ("_input_layer.conv_surface.weight", (192, 112, 1)), ("_input_layer.conv_surface.bias", (192,)),
           ("downsample.linear.weight", (384, 768)), ("downsample.norm.weight", (768,)),
           ("downsample.norm.bias", (768,))]
    for li, (depth, heads, C) in enumerate(zip(DEPTHS, HEADS, DIMS)):
        types = 124 if C == 192 else 64
        for b in range(depth):
            pre = f"layers.EarthSpecificLayer{li}.blocks.EarthSpecificBlock{b}."
            out += [(pre + "norm1.weight", (C,)), (pre + "norm1.bias", (C,)),
                    (pre + "norm2.weight", (C,)), (pre + "norm2.bias", (C,)),
                    (pre + "linear.linear1.weight", (4 * C, C)), (pre + "linear.linear1.bias", (4 * C,)),
                    (pre + "linear.linear2.weight", (C, 4 * C)), (pre + "linear.linear2.bias", (C,)),
                    (pre + "attention.earth_specific_bias", (1, types, heads, 144, 144)),
                    (pre + "attention.linear1.weight", (3 * C, C)), (pre + "attention.linear1.bias", (3 * C,)),
                    (pre + "attention.linear2.weight", (C, C)), (pre + "attention.linear2.bias", (C,))]
    out += [("upsample.linear1.weight", (768, 384)), ("upsample.linear2.weight", (192, 192)),
            ("upsample.norm.weight", (192,)), ("upsample.norm.bias", (192,)),
            ("_output_layer.conv.weight", (160, 384, 1)), ("_output_layer.conv.bias", (160,)),
            ("_output_layer.conv_surface.weight", (64, 384, 1)), ("_output_layer.conv_surface.bias", (64,))]
    return out


def reference_like_weights(seed: int = 0) -> Dict[str, Tensor]:
    """Random weights with the reference's init *distributions* (models/pangu_model.py:41-48,
    models/layers.py:314; Conv1d keeps torch's default U(+-1/sqrt(fan_in))) but drawn in
    ``param_shapes`` order from one Generator, so they can be regenerated anywhere from
    the seed without importing the reference."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes():
        if "norm" in name:
            t = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
        elif "conv" in name:
            bound = 1.0 / math.sqrt(shape[1] if len(shape) == 3 else
                                    {192: 192 if "surface" not in name else 112, 160: 384, 64: 384}[shape[0]])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith("bias") and not name.endswith("earth_specific_bias"):
            t = torch.zeros(shape)
        else:
            t = (torch.randn(shape, generator=g) * 0.02).clamp_(-2.0, 2.0)
        p[name] = t
    return p


def stress_weights(seed: int = 0, bias_std: float = 1.0, w_std: float = 0.05) -> Dict[str, Tensor]:
    """A 'stress' initialisation: large earth-specific bias and
    non-trivial LN/bias vectors so that indexing or mask bugs are visible (at the
    reference's default init attention is almost uniform)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes():
        if name.endswith("earth_specific_bias"):
            t = torch.randn(shape, generator=g) * bias_std
        elif "norm" in name and name.endswith("weight"):
            t = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1]
            t = torch.randn(shape, generator=g) * min(w_std, 1.0 / math.sqrt(fan_in))
        p[name] = t
    return p


def synthetic_inputs(seed: int = 1, lat: int = 721, lon: int = 1440, nontrivial_stats: bool = True):
    """SURVEY.md 8d inputs: upper, surface, maps, const_h drawn in this order from
    Generator(seed); statistics either identity or mean~N(0,1), std~U(0.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    upper = torch.randn(1, 5, 13, lat, lon, generator=g)
    surface = torch.randn(1, 4, lat, lon, generator=g)
    maps = torch.randn(1, 3, 4 * ((lat + 3) // 4), lon, generator=g)
    const_h = torch.randn(1, 1, 1, 13, lat, lon, generator=g)
    if nontrivial_stats:
        stats = (torch.randn(4, generator=g), 0.5 + torch.rand(4, generator=g),
                 torch.randn(13, 1, 1, 5, generator=g), 0.5 + torch.rand(13, 1, 1, 5, generator=g))
    else:
        stats = (torch.zeros(4), torch.ones(4), torch.zeros(13, 1, 1, 5), torch.ones(13, 1, 1, 5))
    return upper, surface, stats, maps, const_h


def lon_strip(upper, surface, maps, const_h, lon0: int, width: int):
    """Cut a longitude strip ``[lon0, lon0+width)`` out of full-grid inputs."""
    sl = slice(lon0, lon0 + width)
    return (upper[..., sl].contiguous(), surface[..., sl].contiguous(),
            maps[..., sl].contiguous(), const_h[..., sl].contiguous())
