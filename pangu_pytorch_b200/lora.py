"""LoRA checkpoints of the reference's ``finetune/lora_tune.py`` -> plain PanguModel weights.

The reference wraps the model with the third-party ``peft`` package
(``LoraConfig(r=16, lora_alpha=16, target_modules=<all 67 nn.Linear>, lora_dropout=0.1,
modules_to_save=["_output_layer.conv_surface", "_output_layer.conv"])``, finetune/lora_tune.py:124-139)
and saves ``peft_model.state_dict()`` (models/pangu_sample.py:94-98).  ``peft`` is not vendored
or pinned by the reference and is absent here, so its published semantics are restated:

    y = W x + b + (lora_alpha / r) * B (A x)          (dropout is the identity at inference)

For inference (``inference/test_lora.py``) the adapters can be folded into the dense weights,
``W' = W + (alpha / r) B A``, after which the checkpoint is an ordinary 223-key state_dict and
the forward is the unmodified B200 hot path -- the rank-16 update costs nothing at run time.
``modules_to_save`` entries replace the corresponding base tensors.

Key layouts accepted (peft 0.4 ... 0.13):
    base_model.model.<path>.weight                      | base_model.model.<path>.base_layer.weight
    base_model.model.<path>.lora_A.<adapter>.weight     | base_model.model.<path>.lora_B.<adapter>.weight
    base_model.model.<path>.original_module.<p>         | base_model.model.<path>.modules_to_save.<adapter>.<p>
"""
from __future__ import annotations

import re
from typing import Dict, Optional

import torch


def merge_lora_state_dict(state: Dict[str, torch.Tensor], lora_alpha: float = 16.0, adapter: str = "default",
                          r: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Fold a peft LoRA ``state_dict`` into plain ``PanguModel`` weights (fp32, on the tensors' device)."""
    pre = "base_model.model."
    plain: Dict[str, torch.Tensor] = {}
    lora_a: Dict[str, torch.Tensor] = {}
    lora_b: Dict[str, torch.Tensor] = {}
    saved: Dict[str, torch.Tensor] = {}
    for k, v in state.items():
        name = k[len(pre):] if k.startswith(pre) else k
        m = re.match(rf"(.+)\.lora_([AB])\.{re.escape(adapter)}\.weight$", name)
        if m:
            (lora_a if m.group(2) == "A" else lora_b)[m.group(1)] = v
            continue
        m = re.match(rf"(.+)\.modules_to_save\.{re.escape(adapter)}\.(.+)$", name)
        if m:
            saved[f"{m.group(1)}.{m.group(2)}"] = v
            continue
        if ".lora_" in name or ".modules_to_save." in name:
            continue                                            # other adapters / lora_dropout etc.
        name = name.replace(".base_layer.", ".").replace(".original_module.", ".")
        plain[name] = v
    if set(lora_a) != set(lora_b):
        raise ValueError("LoRA checkpoint has unmatched lora_A / lora_B tensors")
    out = {k: v.clone() for k, v in plain.items()}
    for mod, a in lora_a.items():
        b = lora_b[mod]
        rank = a.shape[0] if r is None else r
        key = mod + ".weight"
        if key not in out:
            raise KeyError(f"LoRA adapter for '{mod}' has no base weight in the checkpoint")
        if a.shape[1] != out[key].shape[1] or b.shape[0] != out[key].shape[0] or b.shape[1] != a.shape[0]:
            raise ValueError(f"LoRA shapes of '{mod}' do not match the base weight")
        out[key] = out[key].float() + (lora_alpha / rank) * (b.float() @ a.float())
    for key, v in saved.items():                                # trained copies win over the frozen originals
        out[key] = v.clone()
    return out


def load_lora_checkpoint(model, path: str, map_location=None, lora_alpha: float = 16.0):
    """``torch.load(path)['model']`` of a LoRA run -> merged weights -> ``load_state_dict(strict=True)``."""
    ckpt = torch.load(path, map_location=map_location)
    state = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    model.load_state_dict(merge_lora_state_dict(state, lora_alpha=lora_alpha), strict=True)
    return model


# ----------------------------------------------------------------------------------------------
# LoRA finetuning on the B200 path (reference finetune/lora_tune.py:124-139)
# ----------------------------------------------------------------------------------------------
import math

from torch import nn


class LoraLinear(nn.Module):
    """Stand-in for peft's ``lora.Linear`` wrapper with the same child names (``base_layer``,
    ``lora_A.<adapter>``, ``lora_B.<adapter>``), so ``state_dict()`` keys match peft's.

    peft would wrap the ``nn.Linear`` children and add the adapter branch inside their ``forward`` -- but on
    this path the children only HOLD parameters (the kernels read the weights directly), so a peft wrap would
    silently do nothing.  Here the adapter is folded into the operand instead: the kernels see
    ``weight = W + (alpha / r) B A`` (re-formed from the current A, B at every step; exact, since the branch
    is linear), and the backward projects the dense weight gradient onto the adapters:
    ``dA = (alpha/r) B^T dW``, ``dB = (alpha/r) dW A^T``.  This is exact for ``lora_dropout == 0`` and in
    eval mode; a non-zero dropout on the adapter input cannot be folded and is rejected in training mode."""

    def __init__(self, base: nn.Linear, r: int = 16, lora_alpha: float = 16.0, lora_dropout: float = 0.0,
                 adapter: str = "default"):
        super().__init__()
        self.base_layer = base
        self.r, self.lora_alpha, self.scaling, self.adapter = r, lora_alpha, lora_alpha / r, adapter
        self.in_features, self.out_features = base.in_features, base.out_features
        self.lora_dropout = nn.ModuleDict({adapter: nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        dev = base.weight.device
        self.lora_A = nn.ModuleDict({adapter: nn.Linear(base.in_features, r, bias=False, device=dev)})
        self.lora_B = nn.ModuleDict({adapter: nn.Linear(r, base.out_features, bias=False, device=dev)})
        nn.init.kaiming_uniform_(self.lora_A[adapter].weight, a=math.sqrt(5))     # peft's default init
        nn.init.zeros_(self.lora_B[adapter].weight)
        for p in base.parameters():
            p.requires_grad_(False)
        self._merged = None          # persistent fp32 W + s B A (plain attribute: not a parameter, not in state_dict)
        self._merged_key = None

    @property
    def A(self) -> torch.Tensor:
        return self.lora_A[self.adapter].weight

    @property
    def B(self) -> torch.Tensor:
        return self.lora_B[self.adapter].weight

    @property
    def weight(self) -> torch.Tensor:
        """Effective dense weight ``W + (alpha/r) B A`` seen by the kernels.  ONE persistent fp32 tensor per module,
        re-formed IN PLACE (so its ``_version`` advances and the 16-bit operand caches of ``engine.Weight16`` /
        ``Weight16T`` notice) whenever W, A or B changed since it was last formed; the cache key is built from the
        version counters and storage of those three tensors, never from a temporary."""
        drop = self.lora_dropout[self.adapter]
        if self.training and isinstance(drop, nn.Dropout) and drop.p > 0:
            raise NotImplementedError("pangu_pytorch_b200: lora_dropout > 0 cannot be folded into the weight; "
                                      "train with lora_dropout=0 (eval / inference accept any value)")
        key = tuple((t.data_ptr(), t._version, str(t.device)) for t in (self.base_layer.weight, self.A, self.B))
        if key != self._merged_key:
            with torch.no_grad():
                if self._merged is None or self._merged.device != self.base_layer.weight.device:
                    self._merged = torch.empty_like(self.base_layer.weight, dtype=torch.float32)
                torch.addmm(self.base_layer.weight.float(), self.B.float(), self.A.float(), alpha=self.scaling, out=self._merged)
            self._merged_key = key
        return self._merged

    def invalidate(self) -> None:
        """Force the merged weight to be re-formed (only needed after writes that bypass the version counter,
        e.g. through ``param.data``)."""
        self._merged_key = None

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x):           # plain-torch semantics of peft (never used by the B200 hot path)
        y = self.base_layer(x)
        return y + self.scaling * self.lora_B[self.adapter](self.lora_A[self.adapter](self.lora_dropout[self.adapter](x)))


def add_lora(model, r: int = 16, lora_alpha: float = 16.0, lora_dropout: float = 0.0,
             modules_to_save=("_output_layer.conv_surface", "_output_layer.conv")):
    """``get_peft_model(model, LoraConfig(r, lora_alpha, target_modules=<every nn.Linear>, lora_dropout,
    modules_to_save))`` of finetune/lora_tune.py:124-139: freezes the base model, puts a rank-``r`` adapter on
    each of the 67 ``nn.Linear`` modules (4 per block + downsample.linear + upsample.linear1/2; the Conv1d
    layers get none) and keeps ``modules_to_save`` fully trainable.  Returns the model (modified in place)."""
    for p in model.parameters():
        p.requires_grad_(False)
    targets = [(name, m) for name, m in model.named_modules() if isinstance(m, nn.Linear)]
    for name, lin in targets:
        parent = model.get_submodule(name.rsplit(".", 1)[0]) if "." in name else model
        setattr(parent, name.rsplit(".", 1)[-1], LoraLinear(lin, r, lora_alpha, lora_dropout))
    for name in modules_to_save:
        for p in model.get_submodule(name).parameters():
            p.requires_grad_(True)
    return model


def peft_state_dict(model, prefix: str = "base_model.model.", adapter: str = "default",
                    modules_to_save=("_output_layer.conv_surface", "_output_layer.conv"),
                    originals: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """``peft_model.state_dict()`` as the reference saves it for a LoRA run (models/pangu_sample.py:94-98), so that
    the checkpoint loads into the reference's ``peft_model`` (inference/test_lora.py:72) as well as back into this
    package (``merge_lora_state_dict``).  peft's ``ModulesToSaveWrapper`` stores every ``modules_to_save`` module twice:
    ``<path>.original_module.<p>`` (the frozen tensor the run started from) and ``<path>.modules_to_save.<adapter>.<p>``
    (the trained copy).  Here the module holds only the trained tensor; ``originals`` may supply the start values
    (``{"<path>.<p>": tensor}``), otherwise the trained values are written to both entries."""
    saved = tuple(modules_to_save)
    out: Dict[str, torch.Tensor] = {}
    for k, v in model.state_dict().items():
        mod = next((m for m in saved if k.startswith(m + ".")), None)
        if mod is None:
            out[prefix + k] = v
            continue
        p = k[len(mod) + 1:]
        orig = v if originals is None or k not in originals else originals[k]
        out[f"{prefix}{mod}.original_module.{p}"] = orig
        out[f"{prefix}{mod}.modules_to_save.{adapter}.{p}"] = v
    return out
