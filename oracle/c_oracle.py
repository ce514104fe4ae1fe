"""ctypes loader for ``symoracle.c`` (TEST INFRASTRUCTURE, see ``oracle/__init__.py``)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import index_oracle as io

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsymoracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "symoracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/libsymoracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.so_class_contract_vec_f64.restype = ctypes.c_double
        _lib.so_class_contract_vec_f64.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
        _lib.so_class_dump_index.restype = ctypes.c_int64
        _lib.so_class_dump_index.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.so_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().so_num_threads())


def contract_all_indices_with_vector(data, rank: int, dim: int, x) -> float:
    """Packed vector contraction, long-double accumulation, OpenMP over the first index value."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    total = 0.0
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue
        v = np.ascontiguousarray(data[cls], dtype=np.float64).reshape(-1)
        mult = np.asarray(cls, dtype=np.int32)
        size = ctypes.c_int64()
        total += lib().so_class_contract_vec_f64(len(cls), mult.ctypes.data, dim, v.ctypes.data, x.ctypes.data,
                                                 float(io.permclass_multiplicity(cls)), ctypes.byref(size))
        assert size.value == v.size, (cls, size.value, v.size)
    return total


def class_repindex(cls, dim: int) -> np.ndarray:
    rank = sum(cls)
    n = io.permclass_size(cls, dim)
    out = np.empty((n, rank), dtype=np.int32)
    mult = np.asarray(cls, dtype=np.int32)
    got = lib().so_class_dump_index(len(cls), mult.ctypes.data, dim, out.ctypes.data)
    assert got == n
    return out
