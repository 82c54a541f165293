"""argtypes/restype declarations for the entry points beyond the two GEMM engines (kept separate so
``_lib.py`` stays a readable mirror of the first half of include/dfgan_b200.h)."""
from __future__ import annotations

import ctypes as C


def declare(l: C.CDLL) -> None:
    for name, (argtypes, restype) in _TABLE.items():
        if hasattr(l, name):
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = restype


_TABLE: dict = {}
