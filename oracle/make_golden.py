"""Generate the committed golden fixtures under tests/golden by importing and running the
UNMODIFIED reference (``/root/reference/models/{layers,pangu_model}.py``) on CPU fp32.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py [--what maps,blocks,full,stress,train]

The reference has no tests or golden vectors of its own, so these
fixtures -- outputs of the reference itself on seeded synthetic weights/inputs -- are
what pins the oracle (``oracle/pangu_oracle.py``) and, through it, the CUDA path.
Weights come from ``pangu_oracle.reference_like_weights`` / ``stress_weights`` and inputs
from ``pangu_oracle.synthetic_inputs`` so that every consumer can regenerate them from
the seed alone; only sampled outputs + norms are stored.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PANGU_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import pangu_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NSAMP = 8192


def sample_positions(numel: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (NSAMP,), generator=g)


def summarize(t: torch.Tensor, seed: int):
    flat = t.detach().reshape(-1).double()
    pos = sample_positions(flat.numel(), seed)
    return {"pos": pos.numpy().astype(np.int64), "val": t.detach().reshape(-1)[pos].numpy().astype(np.float32),
            "l2": np.float64(flat.norm().item()), "mean": np.float64(flat.mean().item()),
            "shape": np.array(t.shape, dtype=np.int64)}


def put(d: dict, name: str, s: dict):
    for k, v in s.items():
        d[f"{name}.{k}"] = v


def import_reference():
    from models import layers as RL          # noqa
    from models import pangu_model as RM     # noqa
    return RL, RM


# --------------------------------------------------------------------------------------
def gen_maps():
    """Integer contracts: window gather map, shift mask, position_index."""
    RL, _ = import_reference()
    out = {}
    for tag, dim, heads, H in (("hi", 192, 6, 181), ("lo", 384, 12, 91)):
        Z, W = 8, 24
        blk = RL.EarthSpecificBlock(dim, 0.0, heads, device="cpu").eval()
        rec = {}

        class Recorder(torch.nn.Module):
            def forward(self, xw, mask):
                rec["xw"], rec["mask"] = xw.clone(), mask
                return xw

        blk.attention = Recorder()
        T = Z * H * W
        x = torch.zeros(1, T, dim)
        x[0, :, 0] = torch.arange(1, T + 1, dtype=torch.float32)
        for roll in (False, True):
            with torch.no_grad():
                blk(x, Z, H, W, roll)
            m = rec["xw"][..., 0].round().long() - 1          # pad (0) -> -1
            out[f"{tag}.src.roll{int(roll)}"] = m.numpy().astype(np.int32)
            mine = O.window_source_index(Z, H, W, roll)
            assert torch.equal(mine, m), f"window map mismatch {tag} roll={roll}"
            if roll:
                mask = rec["mask"]                             # [nLon, types, 144, 144]
                assert all(torch.equal(mask[0], mask[i]) for i in range(mask.shape[0]))
                bits = (mask[0] != 0)
                assert set(mask.unique().tolist()) <= {0.0, -100.0}
                out[f"{tag}.mask_bits"] = np.packbits(bits.numpy().reshape(-1))
                assert torch.equal(O.shift_mask(Z, H), mask[0]), f"mask mismatch {tag}"
    att = RL.EarthAttention3D(192, 6, 0, (2, 6, 12), device="cpu")
    out["position_index"] = att.position_index.numpy().astype(np.int16)
    assert torch.equal(O.position_index(), att.position_index)
    np.savez_compressed(os.path.join(GOLD, "index_maps.npz"), **out)
    print("maps: ok (oracle == reference, exact)")


def gen_keys():
    """state_dict contract: the reference's 223 keys, shapes and dtypes in registration order."""
    import json
    _, RM = import_reference()
    model = RM.PanguModel(device="cpu")
    sd = model.state_dict()
    rows = [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]
    assert [r[0] for r in rows] == [n for n, _ in O.param_shapes()]
    assert [tuple(r[1]) for r in rows] == [tuple(s) for _, s in O.param_shapes()]
    assert len(list(model.buffers())) == 0
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as fh:
        json.dump(rows, fh, indent=0)
    print(f"keys: {len(rows)} entries, {sum(v.numel() for v in sd.values())} parameters")


def gen_blocks():
    """One EarthSpecificBlock per resolution, stress weights, W=24 strip, both roll states;
    also DownSample at W=24 (the only other W-generic module)."""
    RL, _ = import_reference()
    p = O.stress_weights(seed=7)
    out = {}
    for tag, dim, heads, H, pre in (("hi", 192, 6, 181, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
                                    ("lo", 384, 12, 91, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock1.")):
        Z, W = 8, 24
        blk = RL.EarthSpecificBlock(dim, 0.0, heads, device="cpu").eval()
        blk.load_state_dict({k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}, strict=True)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(1, Z * H * W, dim, generator=g)
        for roll in (False, True):
            with torch.no_grad():
                y = blk(x, Z, H, W, roll)
            mine = O.earth_block(x, p, pre, Z, H, W, heads, roll)
            err = ((mine - y).norm() / y.norm()).item()
            print(f"block {tag} roll={roll}: oracle vs reference rel-L2 {err:.3e}")
            assert err < 2e-6
            put(out, f"{tag}.roll{int(roll)}", summarize(y, 100 + int(roll)))
    ds = RL.DownSample(192).eval()
    ds.load_state_dict({k[len("downsample."):]: v for k, v in p.items() if k.startswith("downsample.")})
    g = torch.Generator().manual_seed(12)
    x = torch.randn(1, 8 * 181 * 24, 192, generator=g)
    with torch.no_grad():
        y = ds(x, 8, 181, 24)
    err = ((O.down_sample(x, p, 8, 181, 24) - y).norm() / y.norm()).item()
    print(f"downsample: oracle vs reference rel-L2 {err:.3e}")
    assert err < 2e-6
    put(out, "down", summarize(y, 102))
    np.savez_compressed(os.path.join(GOLD, "blocks.npz"), **out)


def run_reference_full(p, inputs, train=False):
    _, RM = import_reference()
    torch.manual_seed(0)
    model = RM.PanguModel(device="cpu")
    model.load_state_dict(p, strict=True)
    model.eval()
    taps = {}

    def hook(name):
        def f(_m, _i, o):
            taps[name] = o.detach()
        return f

    model._input_layer.register_forward_hook(hook("embed"))
    for i in range(4):
        model.layers[i].register_forward_hook(hook(f"layer{i}"))
    model.downsample.register_forward_hook(hook("down"))
    model.upsample.register_forward_hook(hook("up"))
    upper, surface, stats, maps, const_h = inputs
    t0 = time.time()
    with torch.no_grad():
        ou, os_ = model(upper, surface, stats, maps, const_h)
    print(f"reference full forward: {time.time() - t0:.1f}s")
    return ou, os_, taps, model


def gen_full(kind: str):
    """Full 0.25-degree forward through the reference PanguModel (config 1 of BASELINE.json)."""
    if kind == "full":
        p, wseed = O.reference_like_weights(seed=0), 0
    else:
        p, wseed = O.stress_weights(seed=3, bias_std=1.0), 3
    inputs = O.synthetic_inputs(seed=1, nontrivial_stats=True)
    ou, os_, taps, _ = run_reference_full(p, inputs)
    out = {"weights_seed": np.int64(wseed), "inputs_seed": np.int64(1)}
    for i, name in enumerate(("embed", "layer0", "down", "layer1", "layer2", "up", "layer3")):
        put(out, name, summarize(taps[name], 200 + i))
    put(out, "out_upper", summarize(ou, 210))
    put(out, "out_surface", summarize(os_, 211))
    out["out_upper.var_l2"] = ou[0].double().flatten(1).norm(dim=1).numpy()
    out["out_surface.var_l2"] = os_[0].double().flatten(1).norm(dim=1).numpy()
    # oracle vs reference on the whole tensors, stage by stage
    otaps = {}
    t0 = time.time()
    mu, ms = O.forward(p, *inputs, taps=otaps)
    print(f"oracle full forward: {time.time() - t0:.1f}s")
    for name in ("embed", "layer0", "down", "layer1", "layer2", "up", "layer3"):
        e = ((otaps[name] - taps[name]).norm() / taps[name].norm()).item()
        print(f"  {kind} {name}: oracle vs reference rel-L2 {e:.3e}")
        assert e < 1e-5
    eu = ((mu - ou).flatten(2).norm(dim=2) / ou.flatten(2).norm(dim=2)).max().item()
    es = ((ms - os_).flatten(2).norm(dim=2) / os_.flatten(2).norm(dim=2)).max().item()
    print(f"  {kind} outputs: per-variable rel-L2 upper {eu:.3e} surface {es:.3e}")
    assert eu < 1e-5 and es < 1e-5
    # second step of a rollout, fed with the reference's own de-normalised output (D9)
    np.savez_compressed(os.path.join(GOLD, f"{kind}_forward.npz"), **out)


def gen_rollout(steps: int = 7):
    """7 x 24 h autoregressive rollout of the UNMODIFIED reference at the full 0.25 degree shapes (BASELINE.json
    configs[1]): x_{k+1} = normBackData(model(x_k)) with the output-order statistics; ~7 min."""
    _, RM = import_reference()
    p = O.reference_like_weights(seed=0)
    upper, surface, stats, maps, const_h = O.synthetic_inputs(seed=1, nontrivial_stats=True)
    torch.manual_seed(0)
    model = RM.PanguModel(device="cpu")
    model.load_state_dict(p, strict=True)
    model.eval()
    ostats = O.output_statistics(stats)
    out = {"weights_seed": np.int64(0), "inputs_seed": np.int64(1), "steps": np.int64(steps)}
    for k in range(steps):
        t0 = time.time()
        with torch.no_grad():
            ou, os_ = model(upper, surface, stats, maps, const_h)
        upper, surface = O.norm_back_data(ou, os_, ostats)
        put(out, f"step{k + 1}.upper", summarize(upper, 300 + 2 * k))
        put(out, f"step{k + 1}.surface", summarize(surface, 301 + 2 * k))
        out[f"step{k + 1}.upper.var_l2"] = upper[0].double().flatten(1).norm(dim=1).numpy()
        out[f"step{k + 1}.surface.var_l2"] = surface[0].double().flatten(1).norm(dim=1).numpy()
        print(f"reference rollout step {k + 1}: {time.time() - t0:.1f}s")
    np.savez_compressed(os.path.join(GOLD, "rollout7.npz"), **out)


def mse_seed_loss(ou, os_, tu, ts):
    """Smooth stand-in for the training loss used ONLY to pin the backward at full resolution: the weighted L1 of
    models/pangu_sample.py:61-67 has a sign() seed that flips wherever two forwards differ by rounding, so gradient
    fixtures are generated with 0.5 * weighted MSE (same per-variable weights, same 0.25 surface factor)."""
    wu = torch.tensor(O.UPPER_WEIGHTS).view(1, 5, 1, 1, 1)
    ws = torch.tensor(O.SURFACE_WEIGHTS).view(1, 4, 1, 1)
    return 0.5 * ((ou - tu) ** 2 * wu).mean() + 0.125 * ((os_ - ts) ** 2 * ws).mean()


def gen_train():
    """Gradients of the UNMODIFIED reference at the full 0.25 degree shapes (BASELINE.json configs[3]): eval mode
    (DropPath off), stress weights, autograd through the reference's own checkpointed blocks.  ~4 min, ~20 GB."""
    _, RM = import_reference()
    p = O.stress_weights(seed=3, bias_std=0.5)
    inputs = O.synthetic_inputs(seed=1, nontrivial_stats=True)
    g = torch.Generator().manual_seed(7)
    tu = torch.randn(1, 5, 13, 721, 1440, generator=g)
    ts = torch.randn(1, 4, 721, 1440, generator=g)
    torch.manual_seed(0)
    model = RM.PanguModel(device="cpu")
    model.load_state_dict(p, strict=True)
    model.eval()
    t0 = time.time()
    ou, os_ = model(*inputs)
    loss = mse_seed_loss(ou, os_, tu, ts)
    loss.backward()
    print(f"reference full forward + backward: {time.time() - t0:.1f}s, loss {loss.item():.6f}")
    out = {"weights_seed": np.int64(3), "inputs_seed": np.int64(1), "targets_seed": np.int64(7),
           "loss": np.float64(loss.item())}
    put(out, "out_upper", summarize(ou, 210))
    put(out, "out_surface", summarize(os_, 211))
    for i, (name, prm) in enumerate(model.named_parameters()):
        gflat = prm.grad.detach().reshape(-1)
        gg = torch.Generator().manual_seed(1000 + i)
        pos = torch.randint(0, gflat.numel(), (256,), generator=gg)
        out[f"grad.{name}.pos"] = pos.numpy().astype(np.int64)
        out[f"grad.{name}.val"] = gflat[pos].numpy().astype(np.float32)
        out[f"grad.{name}.l2"] = np.float64(gflat.double().norm().item())
    np.savez_compressed(os.path.join(GOLD, "train_grads.npz"), **out)


def gen_block_grads():
    """Block-level backward, oracle autograd vs reference autograd (W = 24 strip, both resolutions and roll states):
    pins the oracle's gradients, which the strip-size GPU tests compare against."""
    RL, _ = import_reference()
    p = O.stress_weights(seed=7)
    for tag, dim, heads, H, pre in (("hi", 192, 6, 181, "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock1."),
                                    ("lo", 384, 12, 91, "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock1.")):
        Z, W = 8, 24
        blk = RL.EarthSpecificBlock(dim, 0.0, heads, device="cpu").eval()
        sub = {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}
        blk.load_state_dict(sub, strict=True)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(1, Z * H * W, dim, generator=g)
        r = torch.randn(1, Z * H * W, dim, generator=g)
        for roll in (False, True):
            blk.zero_grad()
            xr = x.clone().requires_grad_(True)
            (blk(xr, Z, H, W, roll) * r).sum().backward()
            leaves = {k: v.clone().requires_grad_(True) for k, v in p.items() if k.startswith(pre)}
            xo = x.clone().requires_grad_(True)
            (O.earth_block(xo, leaves, pre, Z, H, W, heads, roll) * r).sum().backward()
            assert xr.grad.norm() > 0 and torch.isfinite(xr.grad).all()
            worst = ((xo.grad - xr.grad).norm() / xr.grad.norm()).item()
            for k, prm in blk.named_parameters():
                assert prm.grad.norm() > 0 and torch.isfinite(prm.grad).all(), k
                e = ((leaves[pre + k].grad - prm.grad).norm() / prm.grad.norm()).item()
                assert e == e, k                       # not NaN
                worst = max(worst, e)
            print(f"block grads {tag} roll={roll}: oracle autograd vs reference autograd, worst rel-L2 {worst:.3e}")
            assert worst < 5e-5


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="maps,keys,blocks,full,stress")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    for w in args.what.split(","):
        t0 = time.time()
        {"maps": gen_maps, "keys": gen_keys, "blocks": gen_blocks, "full": lambda: gen_full("full"),
         "stress": lambda: gen_full("stress"), "train": gen_train, "blockgrads": gen_block_grads,
         "rollout": gen_rollout}[w]()
        print(f"[{w}] done in {time.time() - t0:.1f}s")
