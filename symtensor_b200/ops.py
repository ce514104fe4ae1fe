"""CUDA implementations of the four ``symalg`` ops, registered on the GPU tensor classes exactly the way a
reference backend mixin registers them (``@Cls.implements(symalg.f)``, symtensor/base.py:1057-1063; in-tree
precedent symtensor/decomp_symmtensor.py:1012-1038).  Each function validates arguments with the reference's
error behaviour, then makes ONE call into the C-ABI with raw device pointers on torch's current stream.
"""
from __future__ import annotations

import numpy as np
import torch

from . import symalg
from ._cabi import c_i64, check, lib
from .flat import CudaFlatSymmetricTensor
from .permcls import _NP2TORCH, _TORCH2NP, CudaPermClsSymmetricTensor, _stream_ptr

_WS_BYTES = int(lib.st_contract_vec_workspace_bytes())


def _np_dtype_of(x) -> np.dtype:
    if isinstance(x, torch.Tensor):
        return _TORCH2NP.get(x.dtype, np.dtype("float64"))
    return np.asarray(x).dtype


def _promote(a_dtype: np.dtype, x) -> torch.dtype:
    """NumPy promotion of the operands (fp32 tensor with fp64 x gives fp64 -- SURVEY.md B.2)."""
    xd = _np_dtype_of(x)
    if xd.kind in "iub":
        xd = a_dtype
    res = np.result_type(a_dtype, xd)
    if res not in _NP2TORCH:
        raise TypeError(f"unsupported result dtype {res}")
    return _NP2TORCH[res]


def _all_close_to_zero(x) -> bool:
    if isinstance(x, torch.Tensor):
        return bool(torch.isclose(x, torch.zeros((), dtype=x.dtype, device=x.device)).all().item())
    return bool(np.isclose(np.asarray(x), 0).all())


def contract_vec_device(A, x_dev: torch.Tensor, out: torch.Tensor, ws: torch.Tensor, begin: int = 0, end: int = None,
                        packed: torch.Tensor = None) -> None:
    """Raw launch: partial sum of packed coordinates [begin, end) into the 0-d/1-element ``out``.
    ``packed`` is the local shard starting at coordinate ``begin`` (default: the whole buffer)."""
    packed = A._buf if packed is None else packed
    total = A._buf.numel() if (end is None and packed is A._buf) else end
    fn = lib.st_contract_vec_f64 if packed.dtype == torch.float64 else lib.st_contract_vec_f32
    check(fn(A.layout, A.rank, c_i64(A.dim), packed.data_ptr(), c_i64(begin), c_i64(total), x_dev.data_ptr(),
             out.data_ptr(), ws.data_ptr(), _stream_ptr(packed.device)))


def _contract_all_indices_with_vector(symtensor, x):
    """symtensor/symalg.py:505-527."""
    if len(x) != symtensor.dim:
        raise ValueError("Dimensions of tensor and vector must match; received "
                         f"{symtensor.dim} (tensor) and {len(x)} (vector).")
    if _all_close_to_zero(x):
        return 0
    cls = type(symtensor)
    tdt = _promote(symtensor.dtype, x)
    buf = symtensor._buf if symtensor._buf.dtype == tdt else symtensor._buf.to(tdt)
    if symtensor._host:
        # host buffers: stream the packed data through the GPU in chunks (copies overlap the kernel)
        if not torch.cuda.is_available():
            raise RuntimeError("symtensor_b200 needs a CUDA device; there is no CPU fallback")
        xh = (x.detach().cpu() if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(tdt).contiguous()
        out = torch.zeros(32 if cls.layout == 0 else 1, dtype=tdt)
        fn = lib.st_contract_vec_host_f64 if tdt == torch.float64 else lib.st_contract_vec_host_f32
        check(fn(cls.layout, symtensor.rank, c_i64(symtensor.dim), buf.data_ptr(), c_i64(buf.numel()), xh.data_ptr(),
                 out.data_ptr()))
        return cls.from_packed(0, 1, out)
    dev = symtensor.device
    with torch.cuda.device(dev):
        xd = (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(device=dev, dtype=tdt).contiguous()
        out = torch.zeros(32 if cls.layout == 0 else 1, dtype=tdt, device=dev)  # packed buffer of a rank-0 tensor
        ws = torch.empty(_WS_BYTES // 8, dtype=torch.float64, device=dev)
        fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32
        check(fn(cls.layout, symtensor.rank, c_i64(symtensor.dim), buf.data_ptr(), c_i64(0), c_i64(buf.numel()),
                 xd.data_ptr(), out.data_ptr(), ws.data_ptr(), _stream_ptr(dev)))
    return cls.from_packed(0, 1, out)


for _cls in (CudaPermClsSymmetricTensor, CudaFlatSymmetricTensor):
    _cls.implements(symalg.contract_all_indices_with_vector)(_contract_all_indices_with_vector)
