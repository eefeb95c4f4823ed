"""Host-side logic that needs no GPU: the C ABI exports what include/pangu_b200.h declares,
the module tree reproduces the reference's state_dict layout, the device index maps (compiled
for the host) are bit-exact against the oracle, and the product refuses to run without CUDA."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import pangu_oracle as O
from tests.util import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from pangu_pytorch_b200 import build
    return build.build()


def test_cabi_exports_every_declared_symbol(lib_path):
    from pangu_pytorch_b200 import _lib
    names = _lib.header_symbols()
    assert len(names) >= 14
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pangu_b200.h but not exported"
    assert sorted(set(names) - {"pangu_last_error"}) == sorted(_lib.SIGNATURES)
    lib.pangu_version.restype = ctypes.c_int
    assert lib.pangu_version() >= 100


def test_cabi_has_no_torch_or_python_dependency(lib_path):
    out = subprocess.run(["ldd", lib_path], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out and "c10" not in out


def test_sass_contains_blackwell_instructions(lib_path):
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass


@pytest.fixture(scope="module")
def model():
    import pangu_pytorch_b200 as pb
    return pb.PanguModel(device="cpu")


def test_state_dict_layout_matches_reference(model):
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as fh:
        ref = json.load(fh)
    sd = model.state_dict()
    assert [k for k, _, _ in ref] == list(sd.keys())
    for k, shape, dtype in ref:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert len(list(model.buffers())) == 0
    assert sum(v.numel() for v in sd.values()) == 276_659_936


def test_state_dict_roundtrip_strict(model):
    p = O.reference_like_weights(seed=0)
    model.load_state_dict(p, strict=True)
    back = model.state_dict()
    assert all(torch.equal(back[k], p[k]) for k in p)


def test_module_api_surface(model):
    import pangu_pytorch_b200 as pb
    blk = model.layers[0].blocks[1]
    assert isinstance(blk, pb.EarthSpecificBlock) and blk.type_of_windows == 124
    assert model.layers[1].blocks[0].type_of_windows == 64
    assert isinstance(blk.attention.linear1, torch.nn.Linear) and blk.attention.scale == 32 ** -0.5
    assert torch.equal(blk.attention.position_index, O.position_index())
    assert [type(b.drop_path).__name__ for b in model.layers[0].blocks] == ["_Identity", "DropPath"]
    assert abs(model.layers[3].blocks[1].drop_path.drop_prob - 0.2) < 1e-6       # linspace(0, .2, 16)
    x = torch.zeros(1, 8, 186, 24, 1)
    assert torch.equal(blk.gen_mask(x)[0], O.shift_mask(8, 181))
    nlinear = sum(isinstance(m, torch.nn.Linear) for m in model.modules())
    assert nlinear == 67                                                         # LoRA targets, SURVEY a17


def test_reference_pickle_aliases():
    import pangu_pytorch_b200 as pb
    pb.install_reference_aliases()
    import models.layers as ML
    import models.pangu_model as MP
    assert MP.PanguModel is pb.PanguModel and ML.EarthSpecificBlock is pb.EarthSpecificBlock
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]


def test_no_cpu_fallback(model):
    up, sf, stats, maps, ch = O.synthetic_inputs(seed=1, lat=721, lon=96)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        with torch.no_grad():
            model.eval()(up, sf, stats, maps, ch)
    with pytest.raises(RuntimeError):
        model.layers[0].blocks[0](torch.zeros(1, 8 * 181 * 24, 192), 8, 181, 24, False)


def test_product_does_not_import_oracle():
    import pathlib
    for f in pathlib.Path(ROOT, "pangu_pytorch_b200").rglob("*.py"):
        assert "oracle" not in f.read_text(), f"{f} references the oracle"


@pytest.mark.parametrize("H,W", [(181, 24), (91, 36)])
@pytest.mark.parametrize("roll", [0, 1])
def test_device_index_maps_bit_exact_on_host(tmp_path, H, W, roll):
    exe = tmp_path / "geom"
    subprocess.run(["nvcc", "-O1", "-o", str(exe), os.path.join(ROOT, "tests", "geometry_host.cu")], check=True)
    out = tmp_path / "maps.bin"
    subprocess.run([str(exe), "8", str(H), str(W), str(roll), str(out)], check=True)
    raw = np.fromfile(out, dtype=np.int32)
    src = O.window_source_index(8, H, W, bool(roll)).reshape(-1).numpy()
    Tp, T = src.size, 8 * H * W
    w2t, t2w, up = raw[:Tp], raw[Tp:Tp + T], raw[Tp + T:]
    assert np.array_equal(w2t, src)
    assert np.array_equal(w2t[t2w], np.arange(T))                 # inverse on real tokens
    # up-sample pixel shuffle + crop; here (H, W) is the HIGH-res grid
    if W % 24 == 0:
        H2, W2 = (H + 1) // 2, W // 2
        T2 = 8 * H2 * W2
        r = np.arange(T2)
        w2, h2, z = r % W2, (r // W2) % H2, r // (W2 * H2)
        for grp in range(4):
            h, w = 2 * h2 + grp // 2, 2 * w2 + grp % 2
            want = np.where(h >= H, -1, (z * H + h) * W + w)
            assert np.array_equal(up[grp * T2:(grp + 1) * T2], want)


def test_lora_merge_matches_restated_peft_semantics():
    """finetune/lora_tune.py:124-139 (peft, third party): y = Wx + b + (alpha/r) B(Ax); merged weights
    reproduce it, modules_to_save replace the frozen originals, result is a strict 223-key state_dict."""
    from pangu_pytorch_b200.lora import merge_lora_state_dict
    g = torch.Generator().manual_seed(4)
    base = {"upsample.linear1.weight": torch.randn(768, 384, generator=g) * 0.02,
            "upsample.norm.weight": torch.ones(192),
            "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock0.linear.linear1.weight": torch.randn(768, 192, generator=g) * 0.02,
            "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock0.linear.linear1.bias": torch.randn(768, generator=g),
            "_output_layer.conv.weight": torch.randn(160, 384, 1, generator=g)}
    peft = {}
    adapters = {}
    for k, v in base.items():
        mod, leaf = k.rsplit(".", 1)
        if leaf == "weight" and v.ndim == 2:
            peft[f"base_model.model.{mod}.base_layer.weight"] = v
            a, b = torch.randn(16, v.shape[1], generator=g) * 0.1, torch.randn(v.shape[0], 16, generator=g) * 0.1
            adapters[mod] = (a, b)
            peft[f"base_model.model.{mod}.lora_A.default.weight"] = a
            peft[f"base_model.model.{mod}.lora_B.default.weight"] = b
        elif mod == "_output_layer.conv":
            peft[f"base_model.model.{mod}.original_module.weight"] = v
            peft[f"base_model.model.{mod}.modules_to_save.default.weight"] = v + 1.0
        elif k.endswith("linear1.bias"):
            peft[f"base_model.model.{mod}.base_layer.bias"] = v
        else:
            peft[f"base_model.model.{k}"] = v
    merged = merge_lora_state_dict(peft, lora_alpha=16.0)
    assert sorted(merged) == sorted(base)
    assert torch.equal(merged["_output_layer.conv.weight"], base["_output_layer.conv.weight"] + 1.0)
    assert torch.equal(merged["upsample.norm.weight"], base["upsample.norm.weight"])
    mod = "layers.EarthSpecificLayer0.blocks.EarthSpecificBlock0.linear.linear1"
    x = torch.randn(7, 192, generator=g)
    a, b = adapters[mod]
    want = O.lora_linear(x, base[mod + ".weight"], base[mod + ".bias"], a, b, alpha=16.0, r=16)
    got = torch.nn.functional.linear(x, merged[mod + ".weight"], merged[mod + ".bias"])
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        bad = dict(peft)
        bad.pop(f"base_model.model.{mod}.lora_B.default.weight")
        merge_lora_state_dict(bad)


def test_add_lora_wraps_67_linears_and_round_trips():
    """finetune/lora_tune.py:124-139: adapters on every nn.Linear (67), base frozen, the two output convs trainable;
    state_dict keys follow peft's child names; the peft-style checkpoint merges back to a strict 223-key state_dict
    whose weights equal the effective weights the kernels read; adapter dropout > 0 is refused in training mode."""
    import pangu_pytorch_b200 as pb
    from pangu_pytorch_b200 import lora
    m = pb.PanguModel(device="cpu")
    plain_keys = list(m.state_dict().keys())
    lora.add_lora(m, r=16, lora_alpha=16.0, lora_dropout=0.0)
    wrapped = {n: mod for n, mod in m.named_modules() if isinstance(mod, lora.LoraLinear)}
    assert len(wrapped) == 67
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert len(trainable) == 2 * 67 + 4 and all(("lora_" in n) or n.startswith("_output_layer.") for n in trainable)
    n0 = "layers.EarthSpecificLayer1.blocks.EarthSpecificBlock0.attention.linear1"
    sd = m.state_dict()
    assert {n0 + ".base_layer.weight", n0 + ".base_layer.bias", n0 + ".lora_A.default.weight",
            n0 + ".lora_B.default.weight"} <= set(sd)
    mod = wrapped[n0]
    assert tuple(mod.A.shape) == (16, 384) and tuple(mod.B.shape) == (1152, 16)
    assert torch.equal(mod.weight, mod.base_layer.weight)                      # B starts at zero (peft init)
    with torch.no_grad():
        mod.B.normal_(0, 0.05, generator=torch.Generator().manual_seed(1))
    eff = mod.base_layer.weight + mod.scaling * (mod.B @ mod.A)
    assert torch.allclose(mod.weight, eff, atol=1e-6) and not mod.weight.requires_grad
    # The merged weight is ONE persistent tensor re-formed in place when A / B / W change (round-1 bug: a fresh temporary per
    # access made the 16-bit operand caches, keyed on (data_ptr, _version), return the first step's weights for ever).
    w_first = mod.weight
    ptr, ver = w_first.data_ptr(), w_first._version
    assert mod.weight.data_ptr() == ptr and mod.weight._version == ver               # unchanged adapters: cache hit, no re-form
    opt = torch.optim.SGD([mod.A, mod.B], lr=0.5)
    for step in range(2):                                                              # two optimiser steps: the key must move each time
        mod.A.grad, mod.B.grad = torch.ones_like(mod.A), torch.ones_like(mod.B)
        opt.step()
        w_now = mod.weight
        assert w_now.data_ptr() == ptr and w_now._version > ver, "merged weight must be updated in place"
        ver = w_now._version
        eff = mod.base_layer.weight + mod.scaling * (mod.B @ mod.A)
        assert torch.allclose(w_now, eff.detach(), atol=1e-5), f"stale merged weight after optimiser step {step}"
    mod.B.data.mul_(2.0)                                                                # .data bypasses the version counter ...
    mod.invalidate()                                                                    # ... so such writers must say so
    eff = mod.base_layer.weight + mod.scaling * (mod.B @ mod.A)
    assert torch.allclose(mod.weight, eff.detach(), atol=1e-5)
    peft_sd = lora.peft_state_dict(m)
    # peft's key layout: wrapped linears as base_layer / lora_A / lora_B, modules_to_save stored twice (ModulesToSaveWrapper)
    pre = "base_model.model."
    assert pre + n0 + ".base_layer.weight" in peft_sd and pre + n0 + ".lora_A.default.weight" in peft_sd
    for conv in ("_output_layer.conv", "_output_layer.conv_surface"):
        for pname in ("weight", "bias"):
            assert f"{pre}{conv}.original_module.{pname}" in peft_sd and f"{pre}{conv}.modules_to_save.default.{pname}" in peft_sd
            assert f"{pre}{conv}.{pname}" not in peft_sd
    assert len(peft_sd) == 223 + 2 * 67 + 4
    merged = lora.merge_lora_state_dict(peft_sd)
    assert sorted(merged) == sorted(plain_keys)
    assert torch.allclose(merged[n0 + ".weight"], eff.detach(), atol=1e-6)
    pb.PanguModel(device="cpu").load_state_dict(merged, strict=True)
    # adapter dropout cannot be folded into the operand: refused while training, accepted in eval
    m2 = torch.nn.Linear(8, 8)
    w = lora.LoraLinear(m2, r=4, lora_alpha=4.0, lora_dropout=0.1)
    w.train()
    with pytest.raises(NotImplementedError):
        _ = w.weight
    w.eval()
    assert w.weight.shape == (8, 8)


def test_bench_stdout_carries_only_the_json_line():
    """bench.py's contract is ONE JSON line on stdout; library chatter written to fd 1 (NCCL's version banner on rank 0)
    must end up on stderr."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench._only_the_json_line_on_stdout(); "
            "os.write(1, b'NCCL version banner\\n'); print('{\"ok\": 1}')" % root)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip() == '{"ok": 1}'
    assert "NCCL version banner" in res.stderr
