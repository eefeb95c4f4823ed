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
