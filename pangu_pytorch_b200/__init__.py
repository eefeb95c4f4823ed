"""pangu_pytorch_b200 -- B200-native (sm_100a) implementation of the Pangu-Weather
forward / backward hot path behind the module API of zhaoshan2/pangu-pytorch.

    from pangu_pytorch_b200 import PanguModel
    model = PanguModel(device="cuda").to("cuda").eval()
    model.load_state_dict(torch.load("pangu_weather_24_torch.pth")["model"])
    out_upper, out_surface = model(upper, surface, statistics, maps, const_h)

Training (``model.train()``; the backward is hand-written kernels, see ``training``), data-parallel gradient
mean (``dist``), LoRA (``lora``), autoregressive rollout and CUDA-graph replay (``rollout``), ensemble sharding
(``ensemble``) and on-device evaluation scores (``ops.scores``) live in the sub-modules of the same names.

``install_reference_aliases()`` registers ``models.layers`` / ``models.pangu_model`` in
``sys.modules`` so that whole-module pickles written by the reference
(``torch.save(model)``, models/pangu_sample.py:164) resolve to these classes.
"""
from __future__ import annotations

import sys

from .engine import (free_workspaces, operand_dtype, set_operand_dtype, set_training_operand_dtype,  # noqa: F401
                     training_operand_dtype)
from .models import layers, pangu_model  # noqa: F401
from .models.layers import (DownSample, EarthAttention3D, EarthSpecificBlock, EarthSpecificLayer, Mlp,  # noqa: F401
                            PatchEmbedding_pretrain, PatchRecovery_pretrain, UpSample)
from .models.pangu_model import PanguModel, load_reference_checkpoint  # noqa: F401

__version__ = "0.1.0"


def install_reference_aliases() -> None:
    import types
    pkg = sys.modules.get("models")
    if pkg is None:
        pkg = types.ModuleType("models")
        pkg.__path__ = []
        sys.modules["models"] = pkg
    pkg.layers, pkg.pangu_model = layers, pangu_model
    sys.modules["models.layers"] = layers
    sys.modules["models.pangu_model"] = pangu_model
