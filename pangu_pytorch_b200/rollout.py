"""Autoregressive rollout of the 24 h model (BASELINE.json config 2).

Semantics follow the only rollout loop in the reference
(``inference/inference_singleOutput.py:92-105``: feed the outputs back as the next inputs);
because the torch ``PanguModel`` emits *normalised* fields, ``normBackData``
(``era5_data/utils_data.py:324-330``) sits between steps.  The de-normalisation
runs in place on the device (``pangu_denorm_fields``); nothing leaves HBM between steps.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import engine, ops


def denormalize_(upper: torch.Tensor, surface: torch.Tensor, statistics) -> Tuple[torch.Tensor, torch.Tensor]:
    """In-place ``normBackData`` with the model's input-order statistics
    (surface (4,), (4,), upper (13,1,1,5) x2 -- level axis reversed w.r.t. the data)."""
    dev = upper.device
    s_mean, s_std = engine.f32(statistics[0], dev).reshape(4), engine.f32(statistics[1], dev).reshape(4)
    u_mean, u_std = engine.f32(statistics[2], dev).reshape(13, 5), engine.f32(statistics[3], dev).reshape(13, 5)
    ops.denorm_fields(upper, surface, s_mean, s_std, u_mean, u_std)
    return upper, surface


def rollout(model, upper, surface, statistics, maps, const_h, steps: int,
            keep_on_device: bool = True) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Run ``steps`` chained forecasts; returns the de-normalised (physical-unit) fields of
    every step.  ``steps=7`` with the 24 h model is the 7-day rollout of BASELINE.json."""
    outs = []
    with torch.no_grad():
        for _ in range(steps):
            ou, os_ = model(upper, surface, statistics, maps, const_h)
            upper, surface = denormalize_(ou, os_, statistics)
            outs.append((upper, surface) if keep_on_device else (upper.cpu(), surface.cpu()))
    return outs
