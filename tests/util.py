"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerances (relative L2 against the fp32 oracle), by operand format.  Evidence:
# SURVEY.md D11 / DESIGN.md "Numerics": rounding the GEMM operands of the *reference itself* to
# bf16 moves its outputs by 6-8e-3 (fp16: ~1e-3); the bounds below leave ~2x head-room.
TOL_MODEL = {"bf16": 2.0e-2, "fp16": 3.0e-3}      # full forward, per variable
TOL_BLOCK = {"bf16": 1.0e-2, "fp16": 1.5e-3}      # one module (block / embed / down / up / recover)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


def sampled_rel_l2(t: torch.Tensor, g, key: str) -> float:
    """rel-L2 of tensor ``t`` against the sampled golden values stored under ``key``."""
    pos = torch.from_numpy(g[f"{key}.pos"])
    val = torch.from_numpy(g[f"{key}.val"]).double()
    assert tuple(g[f"{key}.shape"]) == tuple(t.shape), (tuple(g[f"{key}.shape"]), tuple(t.shape))
    mine = t.detach().reshape(-1).cpu()[pos].double()
    return float((mine - val).norm() / val.norm())


def to_device(p: dict, device):
    return {k: v.to(device) for k, v in p.items()}
