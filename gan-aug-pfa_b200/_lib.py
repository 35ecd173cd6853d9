"""ctypes binding of libgap_b200.so (include/gap_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call fails this module
raises — it never routes around the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libgap_b200.so"

ACT_NONE, ACT_LRELU, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4


class ConvGemmArgs(C.Structure):
    """Mirror of ``gap_conv_gemm_args``."""

    _fields_ = [
        ("src", C.c_void_p * 2),
        ("src_c", C.c_int * 2),
        ("src_ld", C.c_int64 * 2),
        ("n", C.c_int), ("ih", C.c_int), ("iw", C.c_int),
        ("gh", C.c_int), ("gw", C.c_int),
        ("n_phase", C.c_int),
        ("taps_h", C.c_int), ("taps_w", C.c_int),
        ("in_stride", C.c_int),
        ("in_off_h", C.c_int * 2), ("in_off_w", C.c_int * 2),
        ("out_stride", C.c_int),
        ("wpk", C.c_void_p),
        ("w_rows", C.c_int),
        ("n_out", C.c_int),
        ("oh", C.c_int), ("ow", C.c_int),
        ("out", C.c_void_p),
        ("out_ld", C.c_int64),
        ("act", C.c_int),
        ("out2", C.c_void_p),
        ("out2_ld", C.c_int64),
        ("act2", C.c_int),
        ("bias", C.c_void_p),
        ("stats", C.c_void_p),
    ]


class WgradArgs(C.Structure):
    """Mirror of ``gap_wgrad_args``."""

    _fields_ = [
        ("mop", C.c_void_p), ("m_c", C.c_int), ("m_ld", C.c_int64),
        ("nop", C.c_void_p), ("n_c", C.c_int), ("n_ld", C.c_int64),
        ("n", C.c_int), ("gh", C.c_int), ("gw", C.c_int), ("nh", C.c_int), ("nw", C.c_int),
        ("taps_h", C.c_int), ("taps_w", C.c_int), ("stride", C.c_int),
        ("off_h", C.c_int), ("off_w", C.c_int),
        ("out", C.c_void_p), ("ld_m", C.c_int64), ("ld_tap", C.c_int64),
    ]


_lib = None

# name -> (restype, argtypes); every symbol include/gap_b200.h declares is listed here, and
# tests/test_abi.py checks the two stay in sync.
_SIGNATURES = {
    "gap_last_error_string": (C.c_char_p, []),
    "gap_version": (C.c_int, []),
    "gap_sm_count": (C.c_int, []),
    "gap_debug_set": (C.c_int, [C.c_char_p, C.c_int]),
    "gap_conv_gemm": (C.c_int, [C.POINTER(ConvGemmArgs), C.c_void_p]),
    "gap_conv_wgrad": (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
}


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python gan-aug-pfa_b200/build.py` "
                "(there is no CPU fallback for the gap_* kernels)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "gap call") -> None:
    if rc != 0:
        msg = lib().gap_last_error_string().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def debug_set(key: str, value: int) -> None:
    lib().gap_debug_set(key.encode(), int(value))
