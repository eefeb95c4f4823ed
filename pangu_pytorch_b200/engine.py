"""Device-side plumbing shared by the modules: operand precision, 16-bit weight caches and
the per-resolution activation workspaces (memory laid out once, reused every step).

HBM layout of one resolution (``GridWorkspace``), T = Z*H*W real tokens, Tp = window-padded:

    x32      [T , C ] fp32   residual stream, natural token order (updated in place)
    x16      [T , C ] 16-bit shadow of x32, natural order   (A operand of Mlp.linear1, ...)
    x16w[r]  [Tp, C ] 16-bit shadow in WINDOW order for roll state r in {0,1}; the +5 latitude
                             pad rows are zeroed once here and never written again
    qkv      [3C/32, Tp', 32] 16-bit, one plane per (q|k|v, head), window order rows, q pre-scaled
    att      [Tp, C ] 16-bit heads merged; the block path uses the first T rows in NATURAL token order
    hidden   [T , 4C] 16-bit GELU(linear1) activations: training tape only (inference keeps them on the SM)
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch

from . import ops

_OPERANDS = os.environ.get("PANGU_B200_OPERANDS", "bf16").lower()
if _OPERANDS not in ("bf16", "fp16"):
    raise ValueError("PANGU_B200_OPERANDS must be 'bf16' or 'fp16'")


def set_operand_dtype(name: str) -> None:
    """'bf16' (default; range-safe) or 'fp16' (3 more mantissa bits: ~8x lower error, see
    DESIGN.md 'Numerics').  Tensor-core rate is identical (tcgen05 kind::f16)."""
    global _OPERANDS
    name = name.lower()
    if name not in ("bf16", "fp16"):
        raise ValueError("operand dtype must be 'bf16' or 'fp16'")
    _OPERANDS = name


def operand_dtype() -> str:
    return _OPERANDS


# Operand format of the TRAINING path (forward on the tape + backward).  Default fp16 with the fixed 2^16 loss scale of
# training.LOSS_SCALE: the gradients of the 16 earth_specific_bias tables (91 % of all parameters) come from
# dS = P o (dP - rowsum(P o dP)), a cancellation that amplifies the rounding of q / k / v / dO; with bf16 operands they are
# 7-15 % off fp32 autograd, with fp16 (8x finer mantissa) <= 1.6e-2, every other gradient <= 1e-2 (tests/test_gpu_training.py).
# 'bf16' remains selectable for checkpoints whose activations exceed the fp16 range.
_TRAIN_OPERANDS = os.environ.get("PANGU_B200_TRAIN_OPERANDS", "fp16").lower()
if _TRAIN_OPERANDS not in ("bf16", "fp16"):
    raise ValueError("PANGU_B200_TRAIN_OPERANDS must be 'bf16' or 'fp16'")


def set_training_operand_dtype(name: str) -> None:
    global _TRAIN_OPERANDS
    name = name.lower()
    if name not in ("bf16", "fp16"):
        raise ValueError("operand dtype must be 'bf16' or 'fp16'")
    _TRAIN_OPERANDS = name


def training_operand_dtype() -> str:
    return _TRAIN_OPERANDS


class operands:
    """``with engine.operands('fp16'): ...`` -- run a region with another operand format and restore the previous one."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        self.saved = operand_dtype()
        set_operand_dtype(self.name)

    def __exit__(self, *exc):
        set_operand_dtype(self.saved)
        return False


def use_fp16() -> bool:
    return _OPERANDS == "fp16"


def geometry(Z: int, H: int, W: int) -> Tuple[int, int, int, int]:
    """(T, Tp, types, nLon) of the window partition (models/layers.py:145-151, 216-221)."""
    Hp = H + 5
    if Z != 8 or Hp % 6 or W % 12:
        raise ValueError(f"unsupported token grid ({Z},{H},{W})")
    types, nlon = (Z // 2) * (Hp // 6), W // 12
    return Z * H * W, nlon * types * 144, types, nlon


class GridWorkspace:
    def __init__(self, device, Z: int, H: int, W: int, C: int, fp16: bool):
        self.Z, self.H, self.W, self.C, self.fp16 = Z, H, W, C, fp16
        self.T, self.Tp, self.types, self.nlon = geometry(Z, H, W)
        h = ops.dtype16(fp16)
        T, Tp = self.T, self.Tp
        self.x32 = torch.empty(T, C, dtype=torch.float32, device=device)
        self.x16 = torch.empty(T, C, dtype=h, device=device)
        self.x16w = [torch.zeros(Tp, C, dtype=h, device=device) for _ in range(2)]
        # head-major: [3*heads planes][Tp rounded up to 128][32] (csrc/attention_tc.cuh)
        self.qkv = torch.empty(3 * C // 32, (Tp + 127) // 128 * 128, 32, dtype=h, device=device)
        self.att = torch.empty(Tp, C, dtype=h, device=device)
        # Mlp hidden activations are never materialised on the inference path (single-kernel Mlp at both resolutions:
        # csrc/mlp_fused.cuh at C = 192, the CTA-pair kernel csrc/mlp_fused2.cuh at C = 384); the training tape supplies its own
        self.hidden = None


_WORKSPACES: Dict[tuple, GridWorkspace] = {}


def workspace(device, Z: int, H: int, W: int, C: int) -> GridWorkspace:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pangu_pytorch_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, Z, H, W, C, use_fp16())
    ws = _WORKSPACES.get(key)
    if ws is None:
        ops.check_device()
        ws = _WORKSPACES[key] = GridWorkspace(torch.device("cuda", idx), Z, H, W, C, use_fp16())
    return ws


def free_workspaces() -> None:
    _WORKSPACES.clear()


class Weight16:
    """Lazily maintained 16-bit copy of an fp32 parameter (re-cast when the parameter changes)."""

    def __init__(self, k_pad=None):
        self.k_pad = k_pad
        self._key = None
        self._val = None

    def get(self, param: torch.Tensor) -> torch.Tensor:
        key = (param.data_ptr(), param._version, use_fp16(), str(param.device))
        if key != self._key:
            self._val = ops.cast16(param, use_fp16(), self.k_pad)
            self._key = key
        return self._val


class Weight16T:
    """Lazily maintained TRANSPOSED 16-bit copy [in, out_pad] of an fp32 (out, in[, 1]) weight: the B operand
    of the weight's data-gradient GEMM (re-cast when the parameter changes)."""

    def __init__(self, rows_pad=None):
        self.rows_pad = rows_pad
        self._key = None
        self._val = None

    def get(self, param: torch.Tensor) -> torch.Tensor:
        key = (param.data_ptr(), param._version, use_fp16(), str(param.device))
        if key != self._key:
            self._val = ops.cast16_t(param, use_fp16(), rows_pad=self.rows_pad)
            self._key = key
        return self._val


def f32(t: torch.Tensor, device) -> torch.Tensor:
    """Contiguous fp32 view/copy of a small tensor on ``device`` (statistics, biases)."""
    return t.detach().to(device=device, dtype=torch.float32).contiguous()
