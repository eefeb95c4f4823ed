"""B200-native drop-in for the reference's ``models/pangu_model.py`` (PanguModel).

Constructor, attribute names and ``forward(input, input_surface, statistics, maps,
const_h) -> (output, output_surface)`` follow the reference (models/pangu_model.py:9-87);
the 223-entry ``state_dict`` is key-for-key identical, so ONNX-converted and fine-tuned
checkpoints load with ``strict=True``.  The forward chains the fused kernels over two
resident activation workspaces (0.25 deg: 8x181x360x192 and 8x91x180x384); no permuted,
padded, rolled or concatenated tensor is ever materialised.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

from .. import engine
from ..engine import workspace
from .layers import (DownSample, EarthSpecificLayer, PatchEmbedding_pretrain, PatchRecovery_pretrain, UpSample,
                     _trunc_normal_)


class PanguModel(nn.Module):
    def __init__(self, depths=[2, 6, 6, 2], num_heads=[6, 12, 12, 6], dims=[192, 384, 384, 192],
                 patch_size=(2, 4, 4), device=None):
        super().__init__()
        self.device = device
        self._input_layer = PatchEmbedding_pretrain(patch_size, dims[0])
        self.downsample = DownSample(dims[0])
        dpr = [x.item() for x in torch.linspace(0, 0.2, sum(depths))]      # models/pangu_model.py:19
        self.num_layers = len(depths)
        layer_list = OrderedDict()
        for i in range(self.num_layers):
            layer_list["EarthSpecificLayer{}".format(i)] = EarthSpecificLayer(
                depth=depths[i], dim=dims[i], drop_path_ratio_list=dpr[sum(depths[:i]):sum(depths[:i + 1])],
                heads=num_heads[i], use_checkpoint=self.training, device=self.device)
        self.layers = nn.Sequential(layer_list)
        self.upsample = UpSample(dims[-2], dims[-1])
        self._output_layer = PatchRecovery_pretrain(dims[-2])
        self.apply(self._init_weights)

    def _init_weights(self, m):
        """models/pangu_model.py:41-48"""
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, input, input_surface, statistics, maps, const_h):
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            # training step (models/pangu_sample.py:52-71): activations are kept on a tape, the backward is
            # the hand-written kernel chain of pangu_pytorch_b200/training.py
            from ..training import PanguTrainFunction
            return PanguTrainFunction.apply(self, input, input_surface, statistics, maps, const_h,
                                            *self.parameters())
        dev = self._input_layer.conv.weight.device
        lat, lon = input_surface.shape[-2], input_surface.shape[-1]
        Z, H, W = 8, (lat + 3) // 4, lon // 4
        hi = workspace(dev, Z, H, W, 192)
        lo = workspace(dev, Z, (H + 1) // 2, W // 2, 384)
        skip16 = self._skip16(hi)
        taps = getattr(self, "_taps", None)      # tests only: dict that receives a copy of the residual stream after each stage
        tap = (lambda name, ws: taps.__setitem__(name, ws.x32.clone().unsqueeze(0))) if taps is not None else (lambda name, ws: None)
        with torch.no_grad():
            self._input_layer._run(input, input_surface, statistics, maps, const_h, hi)   # -> hi.x32, hi.x16w[0]
            tap("embed", hi)
            self.layers[0]._run(hi, -1, skip16)            # skip connection kept as its 16-bit shadow
            tap("layer0", hi)
            self.downsample._run(hi, lo)                   # -> lo.x32, lo.x16w[0]
            tap("down", lo)
            self.layers[1]._run(lo, 0)                     # hands over in window order to layer 2
            tap("layer1", lo)
            self.layers[2]._run(lo, -1)                    # -> lo.x16 (natural) for the up-sampling GEMM
            tap("layer2", lo)
            self.upsample._run(lo, hi)                     # -> hi.x32, hi.x16w[0]
            tap("up", hi)
            self.layers[3]._run(hi, -1)                    # -> hi.x16 (natural)
            tap("layer3", hi)
            return self._output_layer._run(skip16, hi, lat, lon)   # cat(skip, x) folded into the K loop

    def _skip16(self, hi):
        buf = getattr(hi, "_skip16", None)
        if buf is None:
            buf = hi._skip16 = torch.empty_like(hi.x16)
        return buf


def load_reference_checkpoint(model: PanguModel, path: str, map_location=None) -> PanguModel:
    """``torch.load(path)['model']`` -> ``load_state_dict(strict=True)``, the reference's own
    checkpoint format (finetune/finetune_fully.py:115-116, models/pangu_sample.py:94-98)."""
    ckpt = torch.load(path, map_location=map_location)
    state = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    model.load_state_dict(state, strict=True)
    return model
