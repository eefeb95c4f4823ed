"""Checkpoint IO: ONNX initializer import with the reference's name table and transposition rule
(models/onnx2torch.py:24-52), and the opt-in compressed earth-specific bias (models/layers.py:319-357).  CPU only."""
import os
import struct

import numpy as np
import pytest
import torch

from pangu_pytorch_b200 import checkpoint as ck


# ---- a minimal ONNX (protobuf) writer, test infrastructure: ModelProto{graph{initializer*}}
def _vi(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _ld(fno, payload):
    return _vi((fno << 3) | 2) + _vi(len(payload)) + payload


def _tensor_proto(name, arr, how):
    msg = b""
    if how == "packed_dims":
        msg += _ld(1, b"".join(_vi(d) for d in arr.shape))
    else:
        msg += b"".join(_vi((1 << 3) | 0) + _vi(d) for d in arr.shape)
    dt = {np.dtype("float32"): 1, np.dtype("float16"): 10, np.dtype("int64"): 7}[arr.dtype]
    msg += _vi((2 << 3) | 0) + _vi(dt)
    if how == "float_data":
        msg += _ld(4, arr.astype("<f4").tobytes())
    else:
        msg += _ld(9, arr.tobytes())
    msg += _ld(8, name.encode())
    return msg


def _write_onnx(path, tensors):
    graph = b"".join(_ld(5, _tensor_proto(n, a, how)) for n, a, how in tensors)
    graph += _ld(2, b"synthetic")                                  # GraphProto.name
    model = _vi((1 << 3) | 0) + _vi(8) + _ld(2, b"test") + _ld(7, graph) + _ld(8, _ld(1, b"") + _vi((2 << 3) | 0) + _vi(17))
    with open(path, "wb") as fh:
        fh.write(model)


def test_onnx_initializer_import_follows_the_reference_rule(tmp_path):
    g = torch.Generator().manual_seed(0)
    like = {                                                       # one tensor of every rank the reference handles (:36-52)
        "norm.weight": torch.randn(192, generator=g),
        "linear.weight": torch.randn(576, 192, generator=g),       # 2-D: stored transposed in the ONNX MatMul
        "conv.weight": torch.randn(160, 384, 1, generator=g),
        "attention.earth_specific_bias": torch.randn(1, 4, 2, 144, 144, generator=g),
        "not.in.table": torch.randn(7, generator=g),
    }
    table = [("norm.weight", "b1.a14.weight"), ("linear.weight", "onnx::MatMul_8946"), ("conv.weight", "b1.a1.weight"),
             ("attention.earth_specific_bias", "onnx::Add_8950")]
    src = {k: torch.randn(v.shape, generator=g) for k, v in like.items()}
    tensors = [("b1.a14.weight", src["norm.weight"].numpy(), "float_data"),
               ("onnx::MatMul_8946", np.ascontiguousarray(src["linear.weight"].numpy().T), "raw"),
               ("b1.a1.weight", src["conv.weight"].numpy().astype(np.float16), "packed_dims"),
               ("onnx::Add_8950", src["attention.earth_specific_bias"].numpy(), "packed_dims"),
               ("unrelated_shape_constant", np.array([1, 2, 3], dtype=np.int64), "raw")]
    path = os.path.join(tmp_path, "m.onnx")
    _write_onnx(path, tensors)
    init = ck.read_onnx_initializers(path)
    assert set(init) == {t[0] for t in tensors} and init["unrelated_shape_constant"].tolist() == [1, 2, 3]
    sd, missing = ck.onnx_to_state_dict(path, table, like)
    assert missing == ["not.in.table"] and torch.equal(sd["not.in.table"], like["not.in.table"])
    assert torch.equal(sd["norm.weight"], src["norm.weight"])
    assert torch.equal(sd["linear.weight"], src["linear.weight"])                       # transposed back to (out, in)
    assert torch.equal(sd["conv.weight"], src["conv.weight"].half().float())            # fp16 initializer widened
    assert torch.equal(sd["attention.earth_specific_bias"], src["attention.earth_specific_bias"])
    bad = [("linear.weight", "b1.a14.weight")]
    with pytest.raises(ValueError):
        ck.onnx_to_state_dict(path, bad, {"linear.weight": like["linear.weight"]})


def test_key_table_reader(tmp_path):
    p = os.path.join(tmp_path, "keys.csv")
    with open(p, "w") as fh:
        fh.write("torch_name,onnx_name\n_input_layer.conv.weight,b1.a1.weight\nx.y,\n")
    assert ck.read_key_table(p) == [("_input_layer.conv.weight", "b1.a1.weight"), ("x.y", "")]


@pytest.mark.reference
def test_reference_key_table_covers_the_223_parameters():
    """keys_all.csv of the reference names every tensor of the B200 model exactly once (names from the committed
    fixture tests/golden/state_dict_keys.json, itself generated from the reference's state_dict)."""
    import json
    ref = os.environ.get("PANGU_REFERENCE", "/root/reference")
    table = ck.read_key_table(os.path.join(ref, "keys_all.csv"))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "state_dict_keys.json")) as fh:
        names = [row[0] for row in json.load(fh)]
    assert len(names) == 223 and sorted(t for t, _ in table) == sorted(names)
    assert all(o for _, o in table)


def test_compressed_bias_round_trip():
    g = torch.Generator().manual_seed(1)
    idx = ck.position_index()
    assert idx.shape == (144 * 144,) and int(idx.min()) == 0 and int(idx.max()) == ck.TABLE_ROWS - 1
    assert len(torch.unique(idx)) == ck.TABLE_ROWS
    table = torch.randn(ck.TABLE_ROWS, 5, 3, generator=g)
    full = ck.expand_bias(table)
    assert full.shape == (1, 5, 3, 144, 144)
    # the gather the paper's model performs: bias[t, h, i, j] = table[position_index[i * 144 + j], t, h]
    assert torch.equal(full[0, 2, 1].reshape(-1), table[idx, 2, 1])
    assert torch.equal(ck.compress_bias(full), table)
    sd = {"a.attention.earth_specific_bias": full, "a.norm1.weight": torch.ones(4)}
    small = ck.compress_state_dict(sd)
    assert set(small) == {"a.attention.earth_specific_bias_table", "a.norm1.weight"}
    assert small["a.attention.earth_specific_bias_table"].numel() * 6 < full.numel()
    back = ck.expand_state_dict(small)
    assert set(back) == set(sd) and torch.equal(back["a.attention.earth_specific_bias"], full)
    dense = full + 0.01 * torch.randn(full.shape, generator=g)                          # a finetuned table is no longer compressible
    with pytest.raises(ValueError):
        ck.compress_bias(dense)
    assert torch.allclose(ck.expand_bias(ck.compress_bias(dense, atol=1.0)), full, atol=0.1)
