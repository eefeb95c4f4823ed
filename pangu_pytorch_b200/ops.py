"""Tensor-level wrappers over the C ABI (``include/pangu_b200.h``).

Every function takes CUDA tensors that the caller owns, validates device / dtype /
contiguity, and enqueues the kernels on torch's current stream.  Nothing here computes
with PyTorch ops: torch is used for memory and streams only.
"""
from __future__ import annotations

from ctypes import c_void_p
from typing import Optional

import torch

from . import _lib

Tensor = torch.Tensor
_LAUNCHES = [0]          # kernels launched through this module (bench.py reports it)

# kernels per entry point (for the gpu_launches bookkeeping)
_KERNELS_PER_CALL = {
    "pangu_cast16": 1, "pangu_to_window16": 1, "pangu_patch_embed": 3, "pangu_qkv": 1,
    "pangu_window_attention": 1, "pangu_proj_ln_residual": 1, "pangu_mlp_ln_residual": 2,
    "pangu_downsample": 2, "pangu_upsample": 2, "pangu_patch_recover": 2, "pangu_linear": 1, "pangu_denorm_fields": 1, "pangu_l1_loss": 2,
    "pangu_cast16_t": 1, "pangu_dgrad": 1, "pangu_wgrad": 1, "pangu_colsum16": 1, "pangu_layernorm_bwd": 1, "pangu_gelu_bwd": 1,
    "pangu_window_attention_bwd": 1, "pangu_recover_grad_gather": 1, "pangu_scores": 2,
}


def launches() -> int:
    return _LAUNCHES[0]


def dtype16(fp16: bool) -> torch.dtype:
    return torch.float16 if fp16 else torch.bfloat16


def _p(t: Optional[Tensor], dtype=None, name: str = "tensor"):
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (the B200 path has no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if _DEVICE[0] is None:
        _DEVICE[0] = t.device.index
    elif _DEVICE[0] != t.device.index:
        raise ValueError(f"{name} is on cuda:{t.device.index}, other operands of the call on cuda:{_DEVICE[0]}")
    return c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class EventProfile:
    """Optional per-entry-point device timing (CUDA events on the launching stream).  Used by
    bench.py to time the dominant kernel *inside* the timed region; off by default."""

    def __init__(self):
        self.records = []          # (name, tag, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, tag, a, b in self.records:
            key = f"{name}[{tag}]" if tag else name
            ms, n = out.get(key, (0.0, 0))
            out[key] = (ms + a.elapsed_time(b), n + 1)
        return out


_PROFILE = [None]
_TAG = [""]


def set_profile(p) -> None:
    _PROFILE[0] = p


def set_tag(tag: str) -> None:
    """Label subsequent calls (e.g. 'hi'/'lo') in the event profile."""
    _TAG[0] = tag


_DEVICE = [None]         # device of the tensors of the call being marshalled (set by _p)


def _call(name, *args, kernels=None):
    # The library launches on the CURRENT CUDA device: make sure that is the device the tensors live on (a model on
    # cuda:1 while cuda:0 is current would otherwise get wrong-device launches).
    dev, _DEVICE[0] = _DEVICE[0], None
    if dev is not None and dev != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return _call_on_current(name, *_restream(args), kernels=kernels)
    return _call_on_current(name, *args, kernels=kernels)


def _restream(args):
    """The stream argument was taken on the previously current device: re-take it on the tensors' device."""
    args = list(args)
    if args and isinstance(args[-1], c_void_p):
        args[-1] = _stream()
    return args


def _call_on_current(name, *args, kernels=None):
    prof = _PROFILE[0]
    if prof is None:
        _lib.call(name, *args)
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call(name, *args)
        b.record()
        prof.records.append((name, _TAG[0], a, b))
    _LAUNCHES[0] += _KERNELS_PER_CALL[name] if kernels is None else kernels


def check_device() -> None:
    _lib.call("pangu_check_device")


def cast16(src: Tensor, fp16: bool, k_pad: Optional[int] = None) -> Tensor:
    """fp32 [rows, k] -> 16-bit [rows, k_pad] (zero padded)."""
    src = src.detach().reshape(src.shape[0], -1).contiguous()
    rows, k = src.shape
    kd = k if k_pad is None else k_pad
    out = torch.empty(rows, kd, dtype=dtype16(fp16), device=src.device)
    _call("pangu_cast16", _p(src, torch.float32, "src"), _p(out), rows, k, kd, int(fp16), _stream())
    return out


def cast_rows(src32: Tensor, out16: Tensor, fp16: bool) -> None:
    """fp32 [M, C] -> existing 16-bit [M, C] buffer."""
    M, C = src32.shape
    _call("pangu_cast16", _p(src32, torch.float32, "src"), _p(out16, dtype16(fp16), "dst"), M, C, C, int(fp16),
          _stream())


def to_window16(x32: Tensor, out: Tensor, Z, H, W, C, roll: int, fp16: bool) -> None:
    _call("pangu_to_window16", _p(x32, torch.float32, "x32"), _p(out, dtype16(fp16), "x16w"), Z, H, W, C, roll,
          int(fp16), _stream())


def patch_embed(upper, surface, s_mean, s_std, u_mean, u_std, maps, const_h, w_u16, b_u, w_s16, b_s,
                ws_a_upper, ws_a_surface, x32, x16w, lat, lon, fp16: bool) -> None:
    f = torch.float32
    h = dtype16(fp16)
    _call("pangu_patch_embed", _p(upper, f, "input"), _p(surface, f, "input_surface"), _p(s_mean, f), _p(s_std, f),
          _p(u_mean, f), _p(u_std, f), _p(maps, f, "maps"), _p(const_h, f, "const_h"), _p(w_u16, h), _p(b_u, f),
          _p(w_s16, h), _p(b_s, f), _p(ws_a_upper, h), _p(ws_a_surface, h), _p(x32, f), _p(x16w, h), lat, lon,
          int(fp16), _stream())


def qkv(x16w, w16, bias, out, Z, H, W, C, fp16: bool) -> None:
    h = dtype16(fp16)
    _call("pangu_qkv", _p(x16w, h, "x16w"), _p(w16, h), _p(bias, torch.float32), _p(out, h), Z, H, W, C, int(fp16),
          _stream())


def window_attention(qkv16, earth_bias, out, Z, H, W, C, heads, roll: bool, fp16: bool, window_order_out: bool = False) -> None:
    """out: [T, C] natural token order (default) or [Tp, C] window order with pad rows (window_order_out)."""
    h = dtype16(fp16)
    _call("pangu_window_attention", _p(qkv16, h, "qkv"), _p(earth_bias, torch.float32, "earth_specific_bias"),
          _p(out, h), Z, H, W, C, heads, int(bool(roll)), int(bool(window_order_out)), int(fp16), _stream())


def proj_ln_residual(att16, w16, bias, gamma, beta, x32, x16, Z, H, W, C, roll: bool, res_scale: float,
                     fp16: bool) -> None:
    h, f = dtype16(fp16), torch.float32
    _call("pangu_proj_ln_residual", _p(att16, h), _p(w16, h), _p(bias, f), _p(gamma, f), _p(beta, f), _p(x32, f),
          _p(x16, h), Z, H, W, C, int(bool(roll)), float(res_scale), int(fp16), _stream())


def mlp_ln_residual(x16_in, w1, b1, w2, b2, gamma, beta, ws_hidden, x32, x16_out, Z, H, W, C, roll_out: int,
                    res_scale: float, fp16: bool) -> None:
    h, f = dtype16(fp16), torch.float32
    _call("pangu_mlp_ln_residual", _p(x16_in, h), _p(w1, h), _p(b1, f), _p(w2, h), _p(b2, f), _p(gamma, f),
          _p(beta, f), _p(ws_hidden, h), _p(x32, f), _p(x16_out, h), Z, H, W, C, int(roll_out), float(res_scale),
          int(fp16), _stream(), kernels=1 if ws_hidden is None else 2)


def downsample(x32_in, gamma, beta, w16, ws_a, x32_out, x16w_out, Z, H, W, C, fp16: bool) -> None:
    h, f = dtype16(fp16), torch.float32
    _call("pangu_downsample", _p(x32_in, f), _p(gamma, f), _p(beta, f), _p(w16, h), _p(ws_a, h), _p(x32_out, f),
          _p(x16w_out, h), Z, H, W, C, int(fp16), _stream())


def upsample(x16_in, w1, gamma, beta, w2, ws_a, x32_out, x16w_out, Z, H, W, C_in, C_out, fp16: bool) -> None:
    h, f = dtype16(fp16), torch.float32
    _call("pangu_upsample", _p(x16_in, h), _p(w1, h), _p(gamma, f), _p(beta, f), _p(w2, h), _p(ws_a, h),
          _p(x32_out, f), _p(x16w_out, h), Z, H, W, C_in, C_out, int(fp16), _stream())


def patch_recover(skip16, x16, w_u16, b_u, w_s16, b_s, out_upper, out_surface, Z, H, W, C, lat, lon,
                  fp16: bool) -> None:
    h, f = dtype16(fp16), torch.float32
    _call("pangu_patch_recover", _p(skip16, h), _p(x16, h), _p(w_u16, h), _p(b_u, f), _p(w_s16, h), _p(b_s, f),
          _p(out_upper, f), _p(out_surface, f), Z, H, W, C, lat, lon, int(fp16), _stream())


def linear(a16, w16, bias, out32, out16, gelu: bool, fp16: bool) -> None:
    """out = a16 @ w16.T + bias (tcgen05 GEMM engine); see ``pangu_linear`` in the header."""
    h = dtype16(fp16)
    M, K = a16.shape
    N = w16.shape[0]
    _call("pangu_linear", _p(a16, h), _p(w16, h), _p(bias, torch.float32), _p(out32, torch.float32),
          _p(out16, h), M, N, K, int(bool(gelu)), int(fp16), _stream())


def denorm_fields(upper, surface, s_mean, s_std, u_mean, u_std) -> None:
    """In-place ``normBackData`` on the model outputs (input-order statistics)."""
    f = torch.float32
    lat, lon = surface.shape[-2], surface.shape[-1]
    _call("pangu_denorm_fields", _p(upper, f, "upper"), _p(surface, f, "surface"), _p(s_mean, f), _p(s_std, f),
          _p(u_mean, f), _p(u_std, f), lat, lon, _stream())


UPPER_WEIGHTS = (3.00, 0.60, 1.50, 0.77, 0.54)      # era5_data/config.py:45
SURFACE_WEIGHTS = (1.50, 0.77, 0.66, 3.00)          # era5_data/config.py:46


def l1_loss(out_upper, out_surface, tgt_upper, tgt_surface, s_mean, s_std, u_mean, u_std, want_grad: bool = False,
            upper_weights=UPPER_WEIGHTS, surface_weights=SURFACE_WEIGHTS):
    """Weighted L1 training loss of the reference (targets in physical units, normalised inside).
    Returns (loss[1] device tensor, grad_upper | None, grad_surface | None)."""
    import ctypes
    f = torch.float32
    dev = out_upper.device
    lat, lon = out_surface.shape[-2], out_surface.shape[-1]
    loss = torch.empty(1, dtype=f, device=dev)
    acc = torch.empty(2, dtype=torch.float64, device=dev)
    gu = torch.empty_like(out_upper) if want_grad else None
    gs = torch.empty_like(out_surface) if want_grad else None
    wu = (ctypes.c_float * 5)(*upper_weights)
    ws = (ctypes.c_float * 4)(*surface_weights)
    _call("pangu_l1_loss", _p(out_upper, f), _p(out_surface, f), _p(tgt_upper, f), _p(tgt_surface, f), _p(s_mean, f),
          _p(s_std, f), _p(u_mean, f), _p(u_std, f), ctypes.cast(wu, c_void_p), ctypes.cast(ws, c_void_p), _p(loss, f),
          _p(acc, torch.float64), _p(gu, f), _p(gs, f), lat, lon, _stream())
    return loss, gu, gs


# ----------------------------------------------------------------------------------------------
# backward pass (include/pangu_b200.h, "Backward pass")
# ----------------------------------------------------------------------------------------------
def cast16_t(src: Tensor, fp16: bool, rows_pad: Optional[int] = None, cols_pad: Optional[int] = None) -> Tensor:
    """fp32 [R, C] -> 16-bit transposed copy [C_pad, R_pad] (zero padded)."""
    src = src.detach().reshape(src.shape[0], -1).contiguous()
    R, C = src.shape
    rp, cp = rows_pad or R, cols_pad or C
    out = torch.empty(cp, rp, dtype=dtype16(fp16), device=src.device)
    _call("pangu_cast16_t", _p(src, torch.float32, "src"), _p(out), R, C, rp, cp, int(fp16), _stream())
    return out


def dgrad(a16: Tensor, wt16: Tensor, kind: int, fp16: bool, out32: Optional[Tensor] = None,
          out16: Optional[Tensor] = None, resid32: Optional[Tensor] = None, bias: Optional[Tensor] = None,
          grid=(8, 1, 12), roll: bool = False) -> None:
    """out[M, N] = a16[M, K] @ wt16[N, K].T (+bias); see ``pangu_dgrad`` for the four output kinds."""
    h, f = dtype16(fp16), torch.float32
    M, K = a16.shape
    N = wt16.shape[0]
    if wt16.shape[1] != K:
        raise ValueError(f"dgrad: reduction extents differ ({K} vs {wt16.shape[1]})")
    Z, H, W = grid
    _call("pangu_dgrad", _p(a16, h, "a16"), _p(wt16, h, "wt16"), _p(bias, f), _p(resid32, f), _p(out32, f), _p(out16, h),
          M, N, K, int(kind), Z, H, W, int(bool(roll)), int(fp16), _stream())


def wgrad(dy16: Tensor, x16: Tensor, dw: Tensor, fp16: bool, n_valid: Optional[int] = None,
          k_valid: Optional[int] = None, k_off: int = 0, alpha: float = 1.0) -> None:
    """dw[n, k_off + k] += alpha * sum_m dy16[m, n] * x16[m, k]; dw: fp32 2-D (row pitch = dw.stride(0))."""
    h = dtype16(fp16)
    M = dy16.shape[0]
    if x16.shape[0] != M:
        raise ValueError("wgrad: row counts differ")
    if dw.dim() != 2 or dw.stride(1) != 1 or dw.dtype != torch.float32 or not dw.is_cuda:
        raise ValueError("wgrad: dw must be a 2-D fp32 CUDA tensor with unit column stride")
    N = dy16.shape[1] if n_valid is None else n_valid
    K = x16.shape[1] if k_valid is None else k_valid
    _call("pangu_wgrad", _p(dy16, h, "dy16"), dy16.stride(0), _p(x16, h, "x16"), x16.stride(0), c_void_p(dw.data_ptr()),
          dw.stride(0), int(k_off), M, N, K, float(alpha), int(fp16), _stream())


def colsum16(src16: Tensor, out: Tensor, fp16: bool, n_valid: Optional[int] = None, alpha: float = 1.0) -> None:
    """out[n] += alpha * sum_m src16[m, n] (bias gradients)."""
    M, N = src16.shape
    _call("pangu_colsum16", _p(src16, dtype16(fp16), "src16"), src16.stride(0), _p(out, torch.float32, "out"), M, N,
          N if n_valid is None else n_valid, float(alpha), int(fp16), _stream())


def layernorm_bwd(y: Tensor, g: Tensor, gamma: Tensor, dgamma: Optional[Tensor], dbeta: Optional[Tensor], rows: int,
                  C: int, mode: int, fp16: bool, dx16: Optional[Tensor] = None, dx32: Optional[Tensor] = None,
                  grid=(8, 1, 12), scale: float = 1.0, palpha: float = 1.0, dbias: Optional[Tensor] = None) -> None:
    """``dbias`` (mode 0): [C] += palpha * sum_rows dx -- the bias gradient of the linear that produced ``y``."""
    f = torch.float32
    Z, H, W = grid
    _call("pangu_layernorm_bwd", _p(y, f, "y"), _p(g, f, "g"), _p(gamma, f, "gamma"), _p(dx16, dtype16(fp16)), _p(dx32, f),
          _p(dgamma, f), _p(dbeta, f), _p(dbias, f), rows, C, int(mode), Z, H, W, float(scale), float(palpha), int(fp16), _stream())


def gelu_bwd(dh16: Tensor, pre16: Tensor, fp16: bool, dbias: Optional[Tensor] = None, alpha: float = 1.0) -> None:
    """dh16 *= gelu'(pre16) in place ([M, N]); ``dbias`` [N] += alpha * column sums of the result."""
    h = dtype16(fp16)
    M, N = dh16.shape
    _call("pangu_gelu_bwd", _p(dh16, h, "dh16"), _p(pre16, h, "pre16"), M, N, _p(dbias, torch.float32), float(alpha), int(fp16),
          _stream())


def window_attention_bwd(qkv16, datt16w, earth_bias, dqkv16, dbias, Z, H, W, C, heads, roll: bool, fp16: bool,
                         palpha: float = 1.0, dbqkv: Optional[Tensor] = None) -> None:
    h, f = dtype16(fp16), torch.float32
    _call("pangu_window_attention_bwd", _p(qkv16, h, "qkv"), _p(datt16w, h, "datt"), _p(earth_bias, f, "earth_specific_bias"),
          _p(dqkv16, h, "dqkv"), _p(dbias, f, "dbias"), _p(dbqkv, f, "dbqkv"), Z, H, W, C, heads, int(bool(roll)), float(palpha), int(fp16), _stream())


def recover_grad_gather(d_upper, d_surface, dy_upper, dy_surface, lat, lon, fp16: bool, scale: float = 1.0) -> None:
    h, f = dtype16(fp16), torch.float32
    _call("pangu_recover_grad_gather", _p(d_upper, f, "d_upper"), _p(d_surface, f, "d_surface"), _p(dy_upper, h),
          _p(dy_surface, h), lat, lon, float(scale), int(fp16), _stream())


# ----------------------------------------------------------------------------------------------
# evaluation scores (reference era5_data/score.py, models/pangu_sample.py:236-270)
# ----------------------------------------------------------------------------------------------
def latitude_weights(num_lat: int, device) -> Tensor:
    """``latitude_weighting_factor_torch`` of era5_data/score.py:88-90 (note the reference's 3.1416)."""
    j = torch.arange(num_lat, dtype=torch.float64)
    c = torch.cos(3.1416 / 180.0 * (90.0 - j * 180.0 / float(num_lat - 1)))
    return (num_lat * c / c.sum()).to(dtype=torch.float32, device=device)


def scores(out_upper, out_surface, tgt_upper, tgt_surface, s_mean, s_std, u_mean, u_std, normalised: bool = True):
    """Latitude-weighted RMSE and ACC per plane: returns (rmse_upper [5,13], rmse_surface [4], acc_upper [5,13],
    acc_surface [4]) device tensors.  ``out_*``: model outputs (normalised=True) or physical fields; targets physical;
    statistics in the input order the model takes."""
    f = torch.float32
    dev = out_upper.device
    lat, lon = out_surface.shape[-2], out_surface.shape[-1]
    w = latitude_weights(lat, dev)
    acc_ws = torch.empty(69 * 4, dtype=torch.float64, device=dev)
    rmse = torch.empty(69, dtype=f, device=dev)
    acc = torch.empty(69, dtype=f, device=dev)
    _call("pangu_scores", _p(out_upper, f, "out_upper"), _p(out_surface, f, "out_surface"), _p(tgt_upper, f, "tgt_upper"),
          _p(tgt_surface, f, "tgt_surface"), _p(s_mean, f), _p(s_std, f), _p(u_mean, f), _p(u_std, f), _p(w, f),
          _p(acc_ws, torch.float64), _p(rmse, f), _p(acc, f), lat, lon, int(bool(normalised)), _stream())
    return rmse[:65].view(5, 13), rmse[65:], acc[:65].view(5, 13), acc[65:]
