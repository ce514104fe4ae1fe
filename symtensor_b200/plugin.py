"""Binding of the CUDA hot path into the REAL ``symtensor`` package (Eike-Flath/symtensor), the way a reference backend
mixin is written: a subclass of ``PermClsTorchSymmetricTensor`` (symtensor/torch_symtensor.py:486) whose four hot ops are
registered through the reference's own decorators

    @Cls.implements(symalg.contract_all_indices_with_vector)        symtensor/base.py:1057-1063, symalg.py:505-527
    @Cls.implements(symalg.contract_all_indices_with_matrix)        symalg.py:475-496
    @Cls.implements(symalg.tensordot)                               symalg.py:427-459
    @Cls.implements_ufunc.outer(symalg.add, subtract, multiply)     symtensor/base.py:259-322, symalg.py:294-316

so that ``symalg.f(A, ...)`` on an instance reaches the sm_100a kernels through NumPy's ``__array_function__`` protocol /
``symalg.ufunc_dispatch`` exactly as it reaches the dense defaults today (a subclass registration shadows them,
symtensor/base.py:682-698).  Everything else -- construction, indexing, iteration, serialization -- is the reference's own
class, untouched: ``_data`` stays the reference's ``{class: torch tensor}`` dict, which the ops upload into ONE packed
device buffer (``CudaPermClsSymmetricTensor``), run through the C-ABI and unpack into a new instance of the same class.

    import symtensor                       # the real package (with its real dependencies)
    from symtensor_b200 import plugin
    B200Tensor = plugin.bind()             # idempotent; returns the subclass
    A = B200Tensor(rank=4, dim=200, data={...})
    symtensor.symalg.contract_all_indices_with_vector(A, x)        # -> CUDA

The CUDA backend computes in float32 / float64.  For any other dtype (the reference also stores ints and bools) an op
hands over to the implementation the parent class has registered -- the reference's own default -- which is how the
reference's registries are meant to compose; nothing of this package computes on the CPU.  ``tests/test_plugin.py`` runs the
reference's own API suite (symtensor/testing/api.py) against the bound class.
"""
from __future__ import annotations

from typing import Dict, Tuple, Union

import numpy as np
import torch

_BOUND = {}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("symtensor_b200.plugin: the bound ops need a CUDA device (there is no CPU fallback in this backend)")


def bind(device: str = "cuda:0"):
    """Create (once) and return ``B200PermClsSymmetricTensor``, a subclass of the reference's
    ``PermClsTorchSymmetricTensor`` with the CUDA implementations registered in ITS registries."""
    if device in _BOUND:
        return _BOUND[device]
    try:
        import symtensor  # noqa: F401  -- the real package; this module never imports a stand-in for it
        from symtensor import symalg
        from symtensor.base import SymmetricTensor
        from symtensor.torch_symtensor import PermClsTorchSymmetricTensor
    except ImportError as e:  # pragma: no cover
        raise ImportError("symtensor_b200.plugin.bind() needs the reference package `symtensor` to be importable") from e

    from . import ops
    from . import symalg as st_symalg
    from .permcls import _NP2TORCH, CudaPermClsSymmetricTensor

    FLOATS = (torch.float32, torch.float64)

    class B200PermClsSymmetricTensor(PermClsTorchSymmetricTensor):
        """``PermClsTorchSymmetricTensor`` whose symmetrized contractions run on a B200 (symtensor_b200)."""
        _data: Dict[Tuple[int, ...], Union[torch.Tensor]]

    Cls = B200PermClsSymmetricTensor
    parent_functions = PermClsTorchSymmetricTensor._HANDLED_FUNCTIONS
    parent_outer = PermClsTorchSymmetricTensor._HANDLED_UFUNCS["outer"]

    def torch_dtype_of(x) -> torch.dtype:
        dt = x.dtype if hasattr(x, "dtype") else np.asarray(x).dtype
        if isinstance(dt, torch.dtype):
            return dt
        return _NP2TORCH.get(np.dtype(dt), torch.int64 if np.dtype(dt).kind in "iub" else torch.float64)

    def is_float(x) -> bool:
        if isinstance(x, (int, float)) and not isinstance(x, bool):
            return True  # weakly typed Python scalar: takes the tensor's dtype
        return torch_dtype_of(x) in FLOATS

    def upload(t) -> CudaPermClsSymmetricTensor:
        """Reference tensor -> packed device buffer.  0-d entries are scalar-compressed classes (expanded), empty entries
        are classes a partial dict left out (zeros)."""
        _require_cuda()
        size_of = {c: s for c, s in zip(*_class_sizes(t.rank, t.dim))}
        data = {}
        for c, v in t._data.items():
            v = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
            if v.ndim > 0 and v.numel() == 0 and size_of.get(tuple(c), 0) != 0:
                continue
            data[tuple(c)] = v
        if not data:
            data = {next(iter(size_of)): torch.zeros((), dtype=torch_dtype_of(t))}
        return CudaPermClsSymmetricTensor(rank=t.rank, dim=t.dim, data=data, dtype=torch_dtype_of(t), device=device)

    def _class_sizes(rank, dim):
        from . import combinatorics as comb
        tab = comb.class_table(rank, dim)
        return tab.classes, tab.sizes

    def download(cls, res: CudaPermClsSymmetricTensor):
        """Packed device result -> instance of the reference class (dict constructor: symmetric by construction)."""
        if res.rank == 0:
            return cls(rank=0, dim=1, data={(): res._buf[0].cpu()})
        host = res._buf.cpu()
        tab = res.class_table
        data = {c: host[o:o + s].clone() for c, s, o in zip(tab.classes, tab.sizes, tab.offsets) if len(c) <= res.dim}
        return cls(rank=res.rank, dim=res.dim, data=data)

    def operand(x):
        """Arguments of tensordot / outer: reference tensors are uploaded, arrays and scalars pass through."""
        if isinstance(x, SymmetricTensor):
            return upload(x)
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
        return x

    def all_float(*xs) -> bool:
        return all(is_float(x) for x in xs)

    @Cls.implements(symalg.contract_all_indices_with_vector)
    def _vec(symtensor, x):
        if not isinstance(symtensor, Cls) or not all_float(symtensor, x):
            return parent_functions[symalg.contract_all_indices_with_vector](symtensor, x)
        if len(x) != symtensor.dim:
            raise ValueError("Dimensions of tensor and vector must match; received "
                             f"{symtensor.dim} (tensor) and {len(x)} (vector).")
        if np.isclose(np.asarray(x), 0).all():
            return 0
        res = st_symalg.contract_all_indices_with_vector(upload(symtensor), np.asarray(x))
        return download(type(symtensor), res)

    @Cls.implements(symalg.contract_all_indices_with_matrix)
    def _mat(symtensor, W):
        if not isinstance(symtensor, Cls) or not all_float(symtensor, W):
            return parent_functions[symalg.contract_all_indices_with_matrix](symtensor, W)
        res = st_symalg.contract_all_indices_with_matrix(upload(symtensor), np.asarray(W))
        return download(type(symtensor), res)

    def _dense_nonsymmetric(x) -> bool:
        """tensordot / outer accept ARBITRARY dense operands in the reference (only the result is symmetrized); the
        packed kernels take symmetric operands, so a non-symmetric array of rank >= 2 goes to the reference default."""
        if isinstance(x, SymmetricTensor) or np.ndim(x) < 2:
            return False
        from symtensor import utils
        return not utils.is_symmetric(np.asarray(x))

    @Cls.implements(symalg.tensordot)
    def _tensordot(a, b, axes=2):
        if not all_float(a, b) or _dense_nonsymmetric(a) or _dense_nonsymmetric(b):
            return parent_functions[symalg.tensordot](a, b, axes)
        cls = symalg.result_array(*(x for x in (a, b) if isinstance(x, SymmetricTensor)))
        res = st_symalg.tensordot(operand(a), operand(b), axes=axes)
        return download(cls if issubclass(cls, Cls) else Cls, res)

    @Cls.implements_ufunc.outer(symalg.add, symalg.subtract, symalg.multiply)
    def _outer(ufunc, a, b, **kwargs):
        mirror = {symalg.add: st_symalg.add, symalg.subtract: st_symalg.subtract, symalg.multiply: st_symalg.multiply}[ufunc]
        if "out" in kwargs or not all_float(a, b) or _dense_nonsymmetric(a) or _dense_nonsymmetric(b):
            return parent_outer[ufunc](a, b, **kwargs)
        dima = a.dim if isinstance(a, SymmetricTensor) else (*np.shape(a), 1)[0]
        dimb = b.dim if isinstance(b, SymmetricTensor) else (*np.shape(b), 1)[0]
        if np.ndim(a) != 0 and np.ndim(b) != 0 and dima != dimb:
            return NotImplemented
        cls = symalg.result_array(*(x for x in (a, b) if isinstance(x, SymmetricTensor)))
        res = mirror.outer(operand(a), operand(b))
        return download(cls if issubclass(cls, Cls) else Cls, res)

    Cls.b200_impls = {"contract_all_indices_with_vector": _vec, "contract_all_indices_with_matrix": _mat,
                      "tensordot": _tensordot, "outer": _outer}
    _BOUND[device] = Cls
    return Cls
