"""ctypes binding of the C-ABI declared in ``include/symtensor_b200.h``.

There is NO fallback: if the shared library is missing the import fails loudly (build it with
``python -m symtensor_b200.build``).  Status codes are mapped to the exception types the reference raises.
"""
from __future__ import annotations

import ctypes
import os

from .build import LIB

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_vp = ctypes.c_void_p

if not os.path.exists(LIB):
    raise ImportError(
        f"symtensor_b200: CUDA library not built ({LIB} missing). Run `python -m symtensor_b200.build` "
        "(needs nvcc); there is no CPU fallback.")

lib = ctypes.CDLL(LIB)

# name -> (restype, argtypes); kept in sync with include/symtensor_b200.h (tests/test_cabi.py checks the header)
SIGNATURES = {
    "st_version": (c_int, []),
    "st_last_error": (ctypes.c_char_p, []),
    "st_num_classes": (c_int, [c_int]),
    "st_class_table": (c_int, [c_int, c_i64, c_i32p, c_i32p, c_i64p, c_i64p, c_i64p]),
    "st_indep_size": (c_int, [c_int, c_i64, c_i64p]),
    "st_host_permcls_rank": (c_int, [c_int, c_i64, c_i32p, c_i32p, c_i64p]),
    "st_host_permcls_unrank": (c_int, [c_int, c_i64, ctypes.c_int32, c_i64, c_i32p]),
    "st_host_flat_rank": (c_int, [c_int, c_i64, c_i32p, c_i64p]),
    "st_host_flat_unrank": (c_int, [c_int, c_i64, c_i64, c_i32p]),
    "st_permcls_unrank": (c_int, [c_int, c_i64, ctypes.c_int32, c_i64, c_i64, c_vp, c_vp]),
    "st_permcls_rank": (c_int, [c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "st_flat_unrank": (c_int, [c_int, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "st_flat_rank": (c_int, [c_int, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "st_contract_vec_workspace_bytes": (c_i64, []),
    "st_contract_vec_f64": (c_int, [c_int, c_int, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "st_contract_vec_f32": (c_int, [c_int, c_int, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "st_contract_vec_ex_f64": (c_int, [c_int, c_int, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_vp]),
    "st_contract_vec_ex_f32": (c_int, [c_int, c_int, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_vp]),
    "st_contract_vec_host_f64": (c_int, [c_int, c_int, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "st_contract_vec_host_f32": (c_int, [c_int, c_int, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "st_permcls_to_flat_f64": (c_int, [c_int, c_i64, c_vp, c_vp, c_vp]),
    "st_permcls_to_flat_f32": (c_int, [c_int, c_i64, c_vp, c_vp, c_vp]),
    "st_flat_to_permcls_f64": (c_int, [c_int, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_flat_to_permcls_f32": (c_int, [c_int, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_pack_dense_f64": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_i64, c_i64, c_int, ctypes.c_double, ctypes.c_double, c_vp, c_vp]),
    "st_pack_dense_f32": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_i64, c_i64, c_int, ctypes.c_double, ctypes.c_double, c_vp, c_vp]),
    "st_unpack_dense_f64": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_vp]),
    "st_unpack_dense_f32": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_vp]),
    "st_outer_f64": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_outer_f32": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_outer_op_f64": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_outer_op_f32": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_outer_vec_workspace_bytes": (c_i64, []),
    "st_outer_vec_f64": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_outer_vec_f32": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "st_tensordot_workspace_bytes": (c_int, [c_int, c_int, c_int, c_i64, c_int, c_i64p]),
    "st_tensordot_is_tiled": (c_int, [c_int, c_int, c_int, c_i64, c_int]),
    "st_tensordot_f64": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "st_tensordot_f32": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "st_tensordot_ranges_f32": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "st_contract_mat_workspace_bytes": (c_int, [c_int, c_i64, c_int, c_i64p]),
    "st_contract_mat_f64": (c_int, [c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "st_contract_mat_f32": (c_int, [c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "st_contract_mat_range_bounds": (c_int, [c_int, c_i64, c_i64, c_i64, c_i64p, c_i64p]),
    "st_contract_mat_range_workspace_bytes": (c_int, [c_int, c_i64, c_i64, c_i64, c_int, c_i64p]),
    "st_contract_mat_range_f64": (c_int, [c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "st_contract_mat_range_f32": (c_int, [c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "st_elementwise_unary_f64": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp]),
    "st_elementwise_unary_f32": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp]),
    "st_elementwise_binary_f64": (c_int, [c_int, c_int, c_int, c_int, c_i64, c_vp, c_vp, ctypes.c_double, c_vp, c_vp]),
    "st_elementwise_binary_f32": (c_int, [c_int, c_int, c_int, c_int, c_i64, c_vp, c_vp, ctypes.c_double, c_vp, c_vp]),
    "st_compare_f64": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_int, c_vp, c_vp, c_vp]),
    "st_compare_f32": (c_int, [c_int, c_int, c_int, c_i64, c_vp, c_vp, c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_int, c_vp, c_vp, c_vp]),
    "st_slice_f64": (c_int, [c_int, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "st_slice_f32": (c_int, [c_int, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "st_set_vec_variant": (c_int, [c_int]),
    "st_set_tuning": (c_int, [ctypes.c_char_p, c_i64]),
    "st_debug_vec_timeline": (c_int, [c_vp, c_i64]),
    "st_debug_permcls_successors": (c_i64, [c_int, c_i64, ctypes.c_int32, c_i64, c_i64, c_vp]),
    "st_debug_rowwalk": (c_i64, [c_int, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "st_debug_rowwalk_device": (c_int, [c_int, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "st_debug_rowwalk_device2": (c_int, [c_int, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "st_debug_sym22_tiles": (c_i64, [c_i64, c_i64, c_i64, c_vp, c_i64]),
    "st_debug_sym22_tiles_ranges": (c_i64, [c_i64, c_int, c_vp, c_vp, c_vp, c_i64]),
    "st_debug_sym22_tiles_stats": (c_int, [c_i64, c_int, c_vp, c_vp, c_vp]),
    "st_launch_count": (c_i64, []),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = header / library mismatch
    _f.restype = _res
    _f.argtypes = _args

ST_OK, ST_ERR_INVALID, ST_ERR_CUDA, ST_ERR_UNSUPPORTED, ST_ERR_OVERFLOW = range(5)
LAYOUT_PERMCLS, LAYOUT_FLAT = 0, 1
OUTER_MULTIPLY, OUTER_ADD, OUTER_SUBTRACT = 0, 1, 2
VEC_OVERLAP = 1
MAX_RANK = 16
CLASS_ALIGN = 32

_EXC = {ST_ERR_INVALID: ValueError, ST_ERR_CUDA: RuntimeError, ST_ERR_UNSUPPORTED: NotImplementedError,
        ST_ERR_OVERFLOW: OverflowError}


def check(rc: int) -> None:
    if rc != ST_OK:
        msg = lib.st_last_error().decode("utf-8", "replace")
        raise _EXC.get(rc, RuntimeError)(f"symtensor_b200: {msg}")
