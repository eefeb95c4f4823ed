"""Checkpoint IO around the hot path: the reference's ONNX -> torch weight import
(``models/onnx2torch.py:24-52`` with the ``keys_all.csv`` name table) and an opt-in compressed
storage of the earth-specific bias (``models/layers.py:319-357``).

The ``onnx`` package is not a dependency of the reference's hot path and is absent here, so the
initialisers are read straight from the protobuf wire format (ModelProto.graph = field 7,
GraphProto.initializer = field 5, TensorProto: dims 1, data_type 2, float_data 4, int64_data 7,
name 8, raw_data 9, double_data 10) -- the published ONNX IR schema, nothing else of the graph is
needed.  Tensors come back as numpy views of the file where the encoding allows it (the 1.1 GB
``pangu_weather_24.onnx`` is mapped, not copied).

Nothing here runs on the GPU; it produces an ordinary 223-key ``state_dict``.
"""
from __future__ import annotations

import csv
import mmap
from typing import Dict, Iterable, List, Tuple

import numpy as np
import torch

# TensorProto.DataType -> numpy
_ONNX_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_,
                10: np.float16, 11: np.float64, 12: np.uint32, 13: np.uint64}


def _varint(buf, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _fields(buf, start: int, end: int):
    """Yield (field number, wire type, value | (payload start, payload end)) of one protobuf message."""
    pos = start
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield fno, wt, v
        elif wt == 1:
            yield fno, wt, (pos, pos + 8)
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            yield fno, wt, (pos, pos + n)
            pos += n
        elif wt == 5:
            yield fno, wt, (pos, pos + 4)
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt} at byte {pos}")


def _packed_varints(buf, a: int, b: int) -> List[int]:
    out = []
    while a < b:
        v, a = _varint(buf, a)
        out.append(v)
    return out


def _tensor(buf, a: int, b: int) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name = 1, ""
    raw = floats = doubles = int64s = None
    external = False
    for fno, wt, v in _fields(buf, a, b):
        if fno == 1:
            dims += _packed_varints(buf, *v) if wt == 2 else [v]
        elif fno == 2:
            dtype = v
        elif fno == 4:
            floats = v if wt == 2 else floats
        elif fno == 7:
            int64s = v if wt == 2 else int64s
        elif fno == 8:
            name = bytes(buf[v[0]:v[1]]).decode()
        elif fno == 9:
            raw = v
        elif fno == 10:
            doubles = v if wt == 2 else doubles
        elif fno == 14 and v == 1:
            external = True
    if external:
        raise ValueError(f"initializer '{name}' uses external data, which this reader does not follow")
    if dtype not in _ONNX_DTYPES:
        raise ValueError(f"initializer '{name}': unsupported ONNX data type {dtype}")
    np_dt = np.dtype(_ONNX_DTYPES[dtype])
    if raw is not None:
        arr = np.frombuffer(buf, dtype=np_dt.newbyteorder("<"), count=(raw[1] - raw[0]) // np_dt.itemsize, offset=raw[0])
    elif floats is not None:
        arr = np.frombuffer(buf, dtype="<f4", count=(floats[1] - floats[0]) // 4, offset=floats[0])
    elif doubles is not None:
        arr = np.frombuffer(buf, dtype="<f8", count=(doubles[1] - doubles[0]) // 8, offset=doubles[0])
    elif int64s is not None:
        arr = np.array(_packed_varints(buf, *int64s), dtype=np.int64)
    else:
        arr = np.zeros(0, dtype=np_dt)
    n = int(np.prod(dims)) if dims else arr.size
    if arr.size != n:
        raise ValueError(f"initializer '{name}': {arr.size} elements for dims {dims}")
    return name, arr.reshape(dims)


def read_onnx_initializers(path: str, names: Iterable[str] | None = None) -> Dict[str, np.ndarray]:
    """``{initializer.name: array}`` of an ONNX file -- what ``models/onnx2torch.py:12-16`` builds with
    ``onnx.numpy_helper.to_array``.  ``names``: only these (default: all).  Arrays are read-only views of the
    memory-mapped file."""
    want = set(names) if names is not None else None
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as fh:
        buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
    view = memoryview(buf)
    for fno, wt, v in _fields(view, 0, len(view)):
        if fno != 7 or wt != 2:                      # ModelProto.graph
            continue
        for gno, gwt, gv in _fields(view, *v):
            if gno != 5 or gwt != 2:                 # GraphProto.initializer
                continue
            name, arr = _tensor(view, *gv)
            if want is None or name in want:
                out[name] = arr
    return out


def read_key_table(path: str) -> List[Tuple[str, str]]:
    """``keys_all.csv`` of the reference (columns ``torch_name, onnx_name``; 223 rows)."""
    with open(path, newline="") as fh:
        rows = list(csv.DictReader(fh))
    return [(r["torch_name"].strip(), (r.get("onnx_name") or "").strip()) for r in rows if (r.get("torch_name") or "").strip()]


def onnx_to_state_dict(onnx_path: str, key_table: str | List[Tuple[str, str]], like: Dict[str, torch.Tensor]):
    """The reference's import rule (``models/onnx2torch.py:27-50``) as a pure function: for every torch name of
    ``like`` (a ``PanguModel.state_dict()``) look up the ONNX initialiser, transpose 2-D MatMul weights
    (``:41-44``), copy 1-/3-/5-D tensors verbatim, check the shape.  Returns (state_dict, names with no ONNX
    record) -- the reference leaves those at their initial values."""
    table = read_key_table(key_table) if isinstance(key_table, str) else list(key_table)
    lut = {t: o for t, o in table}
    init = read_onnx_initializers(onnx_path, [o for _, o in table if o])
    out, missing = {}, []
    for name, ref in like.items():
        o = lut.get(name, "")
        if not o or o not in init:
            missing.append(name)
            out[name] = ref.detach().clone()
            continue
        w = torch.from_numpy(np.array(init[o], dtype=np.float32))        # copies out of the mapping
        if ref.dim() == 2:
            w = w.t().contiguous()
        if tuple(w.shape) != tuple(ref.shape):
            raise ValueError(f"{name}: ONNX initializer '{o}' has shape {tuple(w.shape)}, expected {tuple(ref.shape)}")
        out[name] = w
    return out, missing


def import_onnx_weights(model, onnx_path: str, key_table, freeze: bool = True) -> List[str]:
    """``models/onnx2torch.py`` on a B200 ``PanguModel``: loads the pretrained ONNX weights in place and, like the
    reference (``:39,44,48,52``), clears ``requires_grad`` on every tensor it filled.  Returns the torch names
    that had no ONNX record."""
    sd, missing = onnx_to_state_dict(onnx_path, key_table, model.state_dict())
    model.load_state_dict(sd, strict=True)
    if freeze:
        skip = set(missing)
        for name, p in model.named_parameters():
            if name not in skip:
                p.requires_grad_(False)
    return missing


# ----------------------------------------------------------------------------------------------
# compressed earth-specific bias (opt-in converter; the module API keeps the expanded parameter)
# ----------------------------------------------------------------------------------------------
TABLE_ROWS = (2 * 12 - 1) * 6 * 6 * 2 * 2      # 3 312 distinct (dz, dh, dw) relations of a 2 x 6 x 12 window


def position_index() -> torch.Tensor:
    """[144 * 144] index into the 3 312-row table (``models/layers.py:319-357``, closed form)."""
    wz, wh, ww = 2, 6, 12
    k = torch.arange(wz * wh * ww)
    z, h, w = k // (wh * ww), (k // ww) % wh, k % ww
    idx = (z[:, None] + z[None, :] * wz) * (2 * ww - 1) * wh * wh + (h[:, None] + h[None, :] * wh) * (2 * ww - 1) \
        + (w[:, None] - w[None, :] + ww - 1)
    return idx.reshape(-1)


def expand_bias(table: torch.Tensor) -> torch.Tensor:
    """[3312, types, heads] -> [1, types, heads, 144, 144]: the gather the paper's model does every forward
    (``EarthSpecificBias = bias[position_index]``); the reference stores the result as its parameter."""
    idx = position_index().to(table.device)
    full = table[idx]                                           # [144*144, types, heads]
    return full.reshape(144, 144, table.shape[1], table.shape[2]).permute(2, 3, 0, 1).unsqueeze(0).contiguous()


def compress_bias(full: torch.Tensor, atol: float = 0.0) -> torch.Tensor:
    """[1, types, heads, 144, 144] -> [3312, types, heads] (6.3x smaller).  Lossless only while the tensor still has
    the translation structure of the pretrained weights (entries that share a ``position_index`` are equal); raises
    if they differ by more than ``atol`` -- a finetuned dense table cannot be compressed."""
    _, types, heads, n, _ = full.shape
    flat = full[0].permute(2, 3, 0, 1).reshape(n * n, types, heads)
    idx = position_index().to(full.device)
    table = torch.zeros(TABLE_ROWS, types, heads, dtype=full.dtype, device=full.device)
    table[idx] = flat                                           # any representative of each class
    dev = (table[idx] - flat).abs().max().item()
    if dev > atol:
        raise ValueError(f"bias table is not of compressed form (entries of one relation differ by {dev:.3e})")
    return table


def compress_state_dict(state: Dict[str, torch.Tensor], atol: float = 0.0) -> Dict[str, torch.Tensor]:
    """223-key state_dict -> same with every ``...attention.earth_specific_bias`` replaced by
    ``...attention.earth_specific_bias_table`` (1.1 GB -> 0.26 GB for the pretrained weights)."""
    out = {}
    for k, v in state.items():
        if k.endswith("attention.earth_specific_bias"):
            out[k + "_table"] = compress_bias(v, atol)
        else:
            out[k] = v
    return out


def expand_state_dict(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of ``compress_state_dict``: back to the reference's 223-key layout (``load_state_dict(strict=True)``)."""
    out = {}
    for k, v in state.items():
        if k.endswith("attention.earth_specific_bias_table"):
            out[k[: -len("_table")]] = expand_bias(v)
        else:
            out[k] = v
    return out
