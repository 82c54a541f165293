"""ctypes binding of the C ABI declared in include/dfgan_b200.h.

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``gan-codes_b200/_C/libdfgan_b200.so``.  There is no CPU fallback: if the shared object is missing
the first kernel call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFGAN_LIB") or os.path.join(_HERE, "_C", "libdfgan_b200.so")   # DFGAN_LIB: A/B builds

MAX_SRC = 4
MAX_TAPS = 16


class View4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
                ("N", C.c_int32), ("sw", C.c_int64), ("sh", C.c_int64), ("sn", C.c_int64)]


class Tap(C.Structure):
    _fields_ = [("src", C.c_int32), ("dy", C.c_int32), ("dx", C.c_int32), ("wrow", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("src", View4 * MAX_SRC), ("n_src", C.c_int32), ("n_taps", C.c_int32),
        ("taps", Tap * MAX_TAPS),
        ("wpack", C.c_void_p), ("w_rows", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("out", C.c_void_p), ("out_sw", C.c_int64), ("out_sh", C.c_int64), ("out_sn", C.c_int64),
        ("bias", C.c_void_p), ("scale_dev", C.c_void_p), ("scale_imm", C.c_float),
        ("lrelu_pre", C.c_int32), ("slope", C.c_float),
        ("aux", C.c_void_p), ("addend", C.c_void_p),
        ("add_sw", C.c_int64), ("add_sh", C.c_int64), ("add_sn", C.c_int64),
        ("add_shift", C.c_int32), ("add_scale", C.c_float),
        ("mask_src", C.c_void_p), ("n_tile_hint", C.c_int32), ("max_ctas", C.c_int32),
        ("out2", C.c_void_p), ("post_mode", C.c_int32),
        ("tab_ga", C.c_void_p), ("tab_ba", C.c_void_p), ("tab_gb", C.c_void_p), ("tab_bb", C.c_void_p),
        ("tab_stride", C.c_int64), ("out_nchw", C.c_void_p), ("co_valid", C.c_int32), ("tanh_out", C.c_int32),
        ("wpack_img", C.c_void_p), ("img_rows_per_slot", C.c_int32), ("phase_cout", C.c_int32),
        ("workspace", C.c_void_p), ("ws_bytes", C.c_int64),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("src", View4 * MAX_SRC), ("n_src", C.c_int32), ("n_taps", C.c_int32),
        ("taps", Tap * MAX_TAPS),
        ("g", View4), ("Cin", C.c_int32), ("Cout", C.c_int32),
        ("dw", C.c_void_p), ("s_co", C.c_int64), ("s_ci", C.c_int64), ("s_tap", C.c_int64),
        ("scale_dev", C.c_void_p), ("scale_imm", C.c_float), ("k_splits", C.c_int32),
        ("ci_valid", C.c_int32), ("co_valid", C.c_int32),
    ]


class DfganError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the shared object (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DfganError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the sm_100a kernels)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(l: C.CDLL) -> None:
    l.dfgan_version.restype = C.c_int
    l.dfgan_arch.restype = C.c_char_p
    l.dfgan_launch_count.restype = C.c_uint64
    l.dfgan_conv_run.argtypes = [C.POINTER(ConvDesc), C.c_void_p]
    l.dfgan_conv_run.restype = C.c_int
    l.dfgan_conv_uses_image.argtypes = [C.POINTER(ConvDesc)]
    l.dfgan_conv_uses_image.restype = C.c_int
    l.dfgan_conv_splitk_bytes.argtypes = [C.POINTER(ConvDesc)]
    l.dfgan_conv_splitk_bytes.restype = C.c_int64
    l.dfgan_wgrad_run.argtypes = [C.POINTER(WgradDesc), C.c_void_p]
    l.dfgan_wgrad_run.restype = C.c_int
    l.dfgan_profile_enable.argtypes = [C.c_int]
    l.dfgan_profile_enable.restype = None
    l.dfgan_profile_read.argtypes = [C.POINTER(C.c_double)]
    l.dfgan_profile_read.restype = C.c_int
    from . import _lib_decl  # noqa: WPS433  (remaining entry points)
    _lib_decl.declare(l)


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        names = {-1: "invalid argument", -2: "unsupported shape", -3: "driver entry point / tensor map"}
        raise DfganError(f"{what}: {names.get(rc, rc)}")
    raise DfganError(f"{what}: cudaError {rc}")


def profile_enable(on: bool) -> None:
    lib().dfgan_profile_enable(int(on))


def profile_read() -> dict:
    buf = (C.c_double * 8)()
    check(lib().dfgan_profile_read(buf), "dfgan_profile_read")
    keys = ("ms", "flops", "bytes", "launches")
    return {"conv": dict(zip(keys, buf[0:4])), "wgrad": dict(zip(keys, buf[4:8]))}


def launch_count() -> int:
    return int(lib().dfgan_launch_count())
