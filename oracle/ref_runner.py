"""Runs the staged, unmodified reference ``PanguModel`` (``oracle/_ref``, see ``oracle/build_ref.py``) on CPU.

Test / benchmark infrastructure: only ``tests/`` and ``bench.py``'s CPU legs import this.  The reference is imported
in a private module namespace (its own ``models`` package, the ``timm`` shim) and restored afterwards, so it cannot
shadow ``pangu_pytorch_b200.install_reference_aliases``.
"""
from __future__ import annotations

import importlib
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "pangu_model.py")) and os.path.isfile(os.path.join(REF, "models", "layers.py"))


def load_reference():
    """(models.layers, models.pangu_model) of the staged reference."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (python oracle/build_ref.py in the build container)")
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.") or
                  k == "timm" or k.startswith("timm.")}
    try:
        sys.path[:0] = [os.path.join(HERE, "timm_shim"), REF]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            RL = importlib.import_module("models.layers")
            RM = importlib.import_module("models.pangu_model")
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "timm" or k.startswith("timm.")]:
            sys.modules.pop(k)
        sys.modules.update(saved_mods)
    return RL, RM


class ReferenceForward:
    """The reference's fp32 CPU forward at the full 0.25 degree shapes: ``PanguModel(device='cpu')``
    with ``torch.manual_seed(0)`` (its own init), ``.eval()``, ``torch.no_grad()``, all host cores."""

    def __init__(self, threads: int | None = None):
        import torch
        try:                                  # the GPU arm may have pinned this process next to its GPU
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except OSError:
            pass
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        _, RM = load_reference()
        from oracle import pangu_oracle as O
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.model = RM.PanguModel(device="cpu").eval()
        self.inputs = O.synthetic_inputs(seed=1, lat=721, lon=1440)

    def step(self) -> float:
        """One full forward; returns its wall time in seconds."""
        import torch
        up, sf, stats, maps, ch = self.inputs
        t0 = time.perf_counter()
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.out = self.model(up, sf, stats, maps, ch)
        return time.perf_counter() - t0
