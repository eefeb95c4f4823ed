"""ctypes binding of ``libpangu_b200.so`` (C ABI in ``include/pangu_b200.h``).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpangu_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "pangu_b200.h")

_P, _I, _F = c_void_p, c_int, c_float

# name -> argtypes, mirroring include/pangu_b200.h one to one
SIGNATURES = {
    "pangu_version": [],
    "pangu_check_device": [],
    "pangu_cast16": [_P, _P, _I, _I, _I, _I, _P],
    "pangu_to_window16": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pangu_patch_embed": [_P] * 16 + [_I, _I, _I, _P],
    "pangu_qkv": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pangu_window_attention": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pangu_proj_ln_residual": [_P] * 7 + [_I, _I, _I, _I, _I, _F, _I, _P],
    "pangu_mlp_ln_residual": [_P] * 10 + [_I, _I, _I, _I, _I, _F, _I, _P],
    "pangu_downsample": [_P] * 7 + [_I, _I, _I, _I, _I, _P],
    "pangu_upsample": [_P] * 8 + [_I, _I, _I, _I, _I, _I, _P],
    "pangu_patch_recover": [_P] * 8 + [_I, _I, _I, _I, _I, _I, _I, _P],
    "pangu_linear": [_P] * 5 + [_I, _I, _I, _I, _I, _P],
    "pangu_denorm_fields": [_P] * 6 + [_I, _I, _P],
    "pangu_l1_loss": [_P] * 14 + [_I, _I, _P],
    "pangu_scores": [_P] * 12 + [_I, _I, _I, _P],
    "pangu_cast16_t": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pangu_dgrad": [_P] * 6 + [_I] * 9 + [_P],
    "pangu_wgrad": [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    "pangu_colsum16": [_P, _I, _P, _I, _I, _I, _F, _I, _P],
    "pangu_layernorm_bwd": [_P] * 8 + [_I, _I, _I, _I, _I, _I, _F, _F, _I, _P],
    "pangu_gelu_bwd": [_P, _P, _I, _I, _P, _F, _I, _P],
    "pangu_window_attention_bwd": [_P] * 6 + [_I] * 6 + [_F, _I, _P],
    "pangu_recover_grad_gather": [_P] * 4 + [_I, _I, _F, _I, _P],
}

_lib = None


class PanguLibraryError(RuntimeError):
    pass


def header_symbols() -> "list[str]":
    """Function names declared in include/pangu_b200.h."""
    with open(HEADER) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pangu_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    """Load the shared library (building it is ``__graft_entry__.build()``'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PanguLibraryError(
            f"{LIB_PATH} not found: build it with `python -m pangu_pytorch_b200.build` "
            "(there is no CPU or PyTorch fallback for the CUDA path)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.pangu_last_error.restype = c_char_p
    lib.pangu_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.pangu_last_error().decode(errors="replace")
        raise PanguLibraryError(f"{name} failed ({rc}): {msg}")
