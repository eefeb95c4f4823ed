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


class GraphedStep:
    """One forecast step (forward + in-place ``normBackData``) captured in a CUDA graph over static buffers:
    a replay costs one launch instead of ~90 Python -> ctypes -> cudaLaunchKernelEx round trips, so the host
    never paces the rollout.  The kernels, their order and every address are identical to
    the eager path, so a replay is bit-identical to an eager step."""

    def __init__(self, model, upper, surface, statistics, maps, const_h):
        self.model, self.statistics = model, statistics
        self.upper, self.surface = upper.detach().clone(), surface.detach().clone()
        self.maps, self.const_h = maps, const_h
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():          # warm-up: one-time attribute / cache setup outside capture
            ou, os_ = model(self.upper, self.surface, statistics, maps, const_h)
            denormalize_(ou, os_, statistics)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            ou, os_ = model(self.upper, self.surface, statistics, maps, const_h)
            self.out_upper, self.out_surface = denormalize_(ou, os_, statistics)

    def __call__(self, upper=None, surface=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Run one step on (upper, surface) (default: the previous step's outputs fed back).  The returned tensors
        are the graph's static output buffers: clone them if they must survive the next call."""
        self.upper.copy_(self.out_upper if upper is None else upper)
        self.surface.copy_(self.out_surface if surface is None else surface)
        self.graph.replay()
        return self.out_upper, self.out_surface


def rollout_graphed(model, upper, surface, statistics, maps, const_h, steps: int,
                    keep_on_device: bool = True) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """``rollout`` with the step captured once in a CUDA graph and replayed ``steps`` times."""
    step = GraphedStep(model, upper, surface, statistics, maps, const_h)
    outs = []
    for k in range(steps):
        ou, os_ = step(upper, surface) if k == 0 else step()
        outs.append((ou.clone(), os_.clone()) if keep_on_device else (ou.cpu(), os_.cpu()))
    return outs
