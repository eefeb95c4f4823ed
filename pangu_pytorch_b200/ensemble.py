"""Ensemble inference sharded over the GPUs of one box (BASELINE.json config 3).

Members are independent, weights are replicated, so the path shards with NO data-path
collective: member k runs on rank ``k mod world``.  One process per GPU (``torch.distributed.run``);
only host-side metadata (which member ran where, its summary statistics) is gathered, through
``torch.distributed`` if it is initialised.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch


def member_indices(n_members: int, rank: int, world: int) -> List[int]:
    """Round-robin shard: members k with k mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return list(range(rank, n_members, world))


def perturb(upper: torch.Tensor, surface: torch.Tensor, member: int, scale: float = 0.01):
    """Synthetic perturbed member: x + scale * randn with seed 100 + member."""
    g = torch.Generator(device=upper.device).manual_seed(100 + member)
    du = torch.randn(upper.shape, generator=g, device=upper.device, dtype=upper.dtype)
    ds = torch.randn(surface.shape, generator=g, device=surface.device, dtype=surface.dtype)
    return upper + scale * du, surface + scale * ds


def run_ensemble(model, upper, surface, statistics, maps, const_h, n_members: int, rank: int = 0, world: int = 1,
                 reduce: Optional[Callable] = None) -> Dict[int, object]:
    """Forecast this rank's members.  ``reduce(out_upper, out_surface)`` maps each forecast to what
    should be kept (default: the fields themselves).  Returns {member index: result}."""
    results = {}
    with torch.no_grad():
        for k in member_indices(n_members, rank, world):
            pu, ps = perturb(upper, surface, k)
            ou, os_ = model(pu, ps, statistics, maps, const_h)
            results[k] = (ou, os_) if reduce is None else reduce(ou, os_)
    return results


def gather_metadata(local: Dict[int, object]) -> Dict[int, object]:
    """All-gather small python objects (summaries) across ranks; identity without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return dict(sorted(merged.items()))
