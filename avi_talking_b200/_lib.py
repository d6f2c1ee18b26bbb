"""ctypes binding of libavi_b200.so (the C ABI declared in include/avi_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent, importing an op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVI_B200_LIB: developer override (instrumented builds such as profiles/build_timeline_lib.sh); the product loads the in-tree library
LIB_PATH = os.environ.get("AVI_B200_LIB") or os.path.join(_HERE, "libavi_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "avi_b200.h")

ACT_NONE, ACT_GELU, ACT_RELU, ACT_QUICK_GELU = 0, 1, 2, 3
DT_F32, DT_BF16 = 0, 1

_lib = None


class AviGemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("C", C.c_void_p), ("C2", C.c_void_p),
        ("batch", C.c_int32), ("rows", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("conv_taps", C.c_int32), ("conv_stride", C.c_int32),
        ("a_ld", C.c_int64), ("a_batch_stride", C.c_int64), ("a_rows_alloc", C.c_int64),
        ("c_ld", C.c_int64), ("c_batch_stride", C.c_int64),
        ("res_ld", C.c_int64), ("res_batch_stride", C.c_int64),
        ("a_dtype", C.c_int32), ("c_dtype", C.c_int32), ("act", C.c_int32),
        ("conv_taps_x", C.c_int32), ("conv_row_pitch", C.c_int32),
    ]


class AviDecoderWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "sa_in_w", "sa_in_b", "sa_out_w", "sa_out_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b",
        "ln1_w", "ln1_b", "ln2_w", "ln2_b", "ln3_w", "ln3_b", "fb_w", "fb_b", "pe")]


class AviPriorNet(C.Structure):
    _fields_ = [("layers", C.c_void_p), ("learned_query", C.c_void_p), ("rel_bias", C.c_void_p), ("rotary", C.c_void_p),
                ("norm_g", C.c_void_p), ("project_out_t", C.c_void_p),
                ("dim", C.c_int32), ("depth", C.c_int32), ("heads", C.c_int32), ("dim_head", C.c_int32), ("ff_inner", C.c_int32)]


def declared_symbols() -> list[str]:
    """Every function name include/avi_b200.h declares."""
    with open(HEADER_PATH) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avi_[a-z0-9_]+)\s*\(", src)))


def load(check_symbols: bool = False):
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "avi_talking_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.avi_version.restype = C.c_int
        lib.avi_last_error.restype = C.c_char_p
        lib.avi_launch_count.restype = C.c_int64
        _lib = lib
    if check_symbols:
        missing = [s for s in declared_symbols() if not hasattr(_lib, s)]
        if missing:
            raise RuntimeError(f"libavi_b200.so does not export: {missing}")
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().avi_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(load().avi_launch_count())
