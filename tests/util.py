"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerances (relative L2 against the fp32 oracle / reference), by operand format.  Evidence:
# SURVEY.md D11 / DESIGN.md "Numerics": rounding the GEMM operands of the *reference itself* to
# bf16 moves its outputs by 6-8e-3 (worst variable x level 1.0e-2), fp16 by 0.7-1.1e-3; measured on B200:
# bf16 6.0-7.6e-3, fp16 7.9-9.5e-4.  The bounds sit ~1.3-1.6x above the measured values, so an error that
# doubles fails.  north_star's example tolerance (1e-3 for bf16 compute) is met by the fp16 operand format,
# which runs at the same tensor-core rate; bf16 operands cannot meet it (8 mantissa bits), see DESIGN.md 4.
TOL_MODEL = {"bf16": 1.0e-2, "fp16": 1.5e-3}      # full forward, per variable
TOL_BLOCK = {"bf16": 6.0e-3, "fp16": 9.0e-4}      # one module (block / embed / down / up / recover)
TOL_TAP = {"bf16": 1.5e-2, "fp16": 2.0e-3}        # residual stream after each stage of the full forward (measured: bf16 <= 1.16e-2 on the stress init, 7.5e-3 reference-like)


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
