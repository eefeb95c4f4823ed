"""Diagnostic: per-block rel-L2 of the earth_specific_bias gradients against oracle autograd at two strip
widths.  Rounding noise averages over the longitude windows that share a bias tile (error ~ 1/sqrt(nLon));
an indexing bug would not."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pangu_pytorch_b200 as pb
from oracle import pangu_oracle as O
from tests.util import rel_l2

DEV = "cuda:0"
for fmt, lon in (("bf16", 96), ("fp16", 96)):
    pb.set_operand_dtype(fmt)
    pb.free_workspaces()
    p = O.stress_weights(seed=3, bias_std=0.5)
    model = pb.PanguModel(device=DEV)
    model.load_state_dict(p, strict=True)
    model = model.to(DEV).train()
    for blk in [m for m in model.modules() if hasattr(m, "drop_path")]:
        blk.drop_path.drop_prob = 0.0
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=1, lat=721, lon=lon)
    g = torch.Generator().manual_seed(7)
    tu, ts = torch.randn(1, 5, 13, 721, lon, generator=g), torch.randn(1, 4, 721, lon, generator=g)
    out, out_s = model(up.to(DEV), sf.to(DEV), [s.to(DEV) for s in stats], maps.to(DEV), ch.to(DEV))
    _, ref, (gu, gs) = O.loss_and_grads(p, up, sf, stats, maps, ch, tu, ts)
    torch.autograd.backward((out, out_s), (gu.to(DEV), gs.to(DEV)))
    errs = [(n.replace("layers.EarthSpecificLayer", "L").replace(".blocks.EarthSpecificBlock", "B").replace(".attention.earth_specific_bias", ""),
             round(rel_l2(q.grad, ref[n]), 4)) for n, q in model.named_parameters() if n.endswith("earth_specific_bias")]
    print(f"{fmt} lon={lon} (nLon hi/lo = {lon // 48}/{lon // 96}):", errs)
    from pangu_pytorch_b200 import training
    training.release_tape(model)
    del model
