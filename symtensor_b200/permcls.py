"""``CudaPermClsSymmetricTensor``: the reference's permutation-class format with the data resident in HBM.

Mirror of ``PermClsSymmetricTensor`` (symtensor/permcls_symtensor.py:539-979) /
``PermClsTorchSymmetricTensor`` (symtensor/torch_symtensor.py:486-568): ``_data`` is still
``{class tuple: 1-D tensor}`` in ``_perm_classes`` order, but every entry is a *view* into ONE contiguous
device allocation (class starts padded to 32 elements), so that the CUDA kernels see a single packed range
that can be streamed, sharded by ``[begin, end)`` and handed over by raw pointer.

Differences from the reference, on purpose:
* scalar-compressed classes (0-d entries) are expanded when stored; classes missing from a dict are zeros;
* ``device="host"`` keeps the packed buffer in (pinned) host memory and streams it through the GPU per op --
  this is the end-to-end path with host buffers; all arithmetic still runs in the CUDA kernels;
* there is no dense CPU fallback for any op.
"""
from __future__ import annotations

import ctypes
import itertools
import math
from numbers import Number
from typing import Optional, Dict, Iterable, Tuple

import numpy as np
import torch

from . import combinatorics as comb
from ._cabi import LAYOUT_PERMCLS, c_i64, check, lib
from .base import SymmetricTensor
from .elementwise import PackedElementwise

Cls = Tuple[int, ...]

_NP2TORCH = {np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}


def to_torch_dtype(dtype) -> torch.dtype:
    if dtype is None:
        return torch.float64
    if isinstance(dtype, torch.dtype):
        if dtype not in _TORCH2NP:
            raise TypeError(f"unsupported dtype {dtype}: the CUDA backend stores float32 or float64")
        return dtype
    npd = np.dtype(dtype)
    if npd.kind in "iub":  # integer data is promoted like the reference's float default
        return torch.float64
    if npd not in _NP2TORCH:
        raise TypeError(f"unsupported dtype {npd}: the CUDA backend stores float32 or float64")
    return _NP2TORCH[npd]


def _is_host(device) -> bool:
    return str(device) in ("host", "cpu")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _parse_class_key(key) -> Cls:
    if isinstance(key, str):
        if key and (key[0].isdigit() or key[0] == "("):  # serialised tuple, e.g. "(2, 1)"
            return tuple(int(t) for t in key.replace("(", " ").replace(")", " ").replace(",", " ").split())
        return comb.permclass_label_to_counts(key)
    return tuple(int(k) for k in key)


def pack_dense_device(layout: int, rank: int, dim: int, dense: torch.Tensor, buf: torch.Tensor, symmetrize: bool) -> bool:
    """Fill the packed buffer ``buf`` from the dense ``dim**rank`` device tensor with ``pack_dense_kernel``; returns whether the
    dense array passed the reference's symmetry check (``utils.is_symmetric``, symtensor/utils.py:563-578; always True with
    ``symmetrize``, which averages over the axis permutations instead, symtensor/utils.py:507-532)."""
    fn = lib.st_pack_dense_f64 if buf.dtype == torch.float64 else lib.st_pack_dense_f32
    dense = dense.contiguous()
    flag = torch.zeros(1, dtype=torch.int32, device=buf.device)
    with torch.cuda.device(buf.device):
        check(fn(layout, rank, c_i64(dim), dense.data_ptr(), buf.data_ptr(), c_i64(0), c_i64(buf.numel()), int(bool(symmetrize)),
                 1e-5, 1e-8, flag.data_ptr(), _stream_ptr(buf.device)))
    return symmetrize or int(flag.item()) == 0


def unpack_dense_device(layout: int, rank: int, dim: int, buf: torch.Tensor) -> torch.Tensor:
    """``todense`` with ``unpack_dense_kernel`` (one thread per dense element)."""
    fn = lib.st_unpack_dense_f64 if buf.dtype == torch.float64 else lib.st_unpack_dense_f32
    dense = torch.empty((dim,) * rank, dtype=buf.dtype, device=buf.device)
    with torch.cuda.device(buf.device):
        check(fn(layout, rank, c_i64(dim), buf.data_ptr(), dense.data_ptr(), _stream_ptr(buf.device)))
    return dense


def _kernel_dtype(tdt: torch.dtype) -> bool:
    return tdt in (torch.float32, torch.float64)


def _gpu_for_host() -> Optional[torch.device]:
    """Host-resident tensors run their dense <-> packed conversions on the GPU as well (copy in, kernel, copy out) whenever a
    device exists; the index-gather loops below them are the container logic for machines WITHOUT any CUDA device (the CPU
    test-suite of the host logic) -- no contraction runs there."""
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None


class CudaPermClsSymmetricTensor(PackedElementwise, SymmetricTensor):
    """On creation, defaults to a zero tensor (like the reference)."""

    data_format = "PermCls"
    layout = LAYOUT_PERMCLS
    array_type = torch.Tensor

    def __init__(self, rank=None, dim=None, data=np.float64(0), dtype=None, symmetrize: bool = False, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "host"
        self._host = _is_host(device)
        self.device = torch.device("cpu") if self._host else torch.device(device)
        dense = None
        if isinstance(data, (np.ndarray, torch.Tensor)) and not (isinstance(data, dict)):
            dense = data
            if rank is None:
                rank = dense.ndim
            if dim is None:
                dim = max(tuple(dense.shape) + (1,)) if dense.ndim else 1
            if dtype is None:
                dtype = dense.dtype if isinstance(dense, torch.Tensor) else np.asarray(dense).dtype
        if rank is None or dim is None:
            raise NotImplementedError("rank and dim must be given unless `data` is a dense array")
        self.rank, self.dim = int(rank), int(dim)
        self._table = comb.class_table(self.rank, self.dim)
        if dtype is None and isinstance(data, dict):
            dts = [v.dtype if isinstance(v, torch.Tensor) else np.asarray(v).dtype for v in data.values()]
            dts = [_TORCH2NP.get(d, d) if isinstance(d, torch.dtype) else d for d in dts]
            dtype = np.result_type(*dts) if dts else None
        elif dtype is None and isinstance(data, Number):
            dtype = np.asarray(data).dtype
        self._tdtype = to_torch_dtype(dtype)
        self._alloc()
        if dense is not None:
            self._init_from_dense(dense, symmetrize)
        elif isinstance(data, Number):
            if data != 0:
                for v in self._data.values():
                    v.fill_(data)
        elif isinstance(data, dict):
            self._init_from_dict(data)
        else:
            raise TypeError("If provided, `data` must be a scalar, a dense array or a dictionary with the format "
                            "{σ class: data vector}")

    # ---- storage ------------------------------------------------------------------------------------
    def _alloc(self, buf=None):
        t = self._table
        if buf is None:
            if self._host:
                buf = torch.zeros(t.total, dtype=self._tdtype, pin_memory=torch.cuda.is_available())
            else:
                buf = torch.zeros(t.total, dtype=self._tdtype, device=self.device)
        assert buf.shape == (t.total,) and buf.dtype == self._tdtype
        self._buf = buf
        self._data: Dict[Cls, torch.Tensor] = {}
        for c, size, off in zip(t.classes, t.sizes, t.offsets):
            v = buf[off:off + size]
            self._data[c] = v.reshape(()) if self.rank == 0 else v

    @classmethod
    def from_packed(cls, rank: int, dim: int, buf: torch.Tensor):
        """Wrap an existing packed buffer (length ``class_table(rank, dim).total``) without copying."""
        self = cls.__new__(cls)
        self.rank, self.dim = int(rank), int(dim)
        self._table = comb.class_table(self.rank, self.dim)
        self._tdtype = to_torch_dtype(buf.dtype)
        self._host = buf.device.type == "cpu"
        self.device = buf.device
        self._alloc(buf)
        return self

    def _init_from_dict(self, data: dict):
        if len(data) == 0:
            raise NotImplementedError("Initializating with empty data is not implemented")
        seen = set()
        for key, v in data.items():
            c = _parse_class_key(key)
            if c in seen:
                raise ValueError(f"`data` contains the key '{key}' twice: possibly in both its original and "
                                 "serialized (str) form.")
            seen.add(c)
            if c not in self._data:
                raise ValueError("`data` argument to PermClsSymmetricTensor does not have the expected format.\n"
                                 f"Expected keys to be a subset of: {sorted(self._data)}\nReceived keys:{sorted(seen)}")
            self._assign_class(c, v)

    def _assign_class(self, c: Cls, v):
        dst = self._data[c]
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(np.asarray(v))
        if v.ndim > 0 and self.rank > 0 and tuple(v.shape) != tuple(dst.shape):
            raise ValueError(f"Data for permutation class {comb.permclass_counts_to_label(c)} should have shape "
                             f"{tuple(dst.shape)}, but the provided data has shape {tuple(v.shape)}.")
        dst.copy_(v.to(dtype=self._tdtype) if v.ndim == 0 else v.reshape(dst.shape))

    def _rep_index_tensor(self, c: Cls) -> torch.Tensor:
        """int64 [size, rank] representative multi-indices of class ``c`` in storage order (GPU enumerator)."""
        ci = self._table.index(c)
        size = self._table.sizes[ci]
        if self._host or size == 0 or self.rank == 0:
            idx = [comb.index_of(self.rank, self.dim, c, p) for p in range(size)]
            return torch.tensor(idx, dtype=torch.int64).reshape(size, self.rank).to(self.device)
        out = torch.empty((size, self.rank), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.st_permcls_unrank(self.rank, c_i64(self.dim), ci, c_i64(0), c_i64(size), out.data_ptr(),
                                        _stream_ptr(self.device)))
        return out.to(torch.int64)

    def _init_from_dense(self, dense, symmetrize: bool):
        if not isinstance(dense, torch.Tensor):
            dense = torch.as_tensor(np.asarray(dense))
        shape = (self.dim,) * self.rank
        try:
            dense = torch.broadcast_to(dense.to(self.device, self._tdtype), shape)
        except RuntimeError as e:
            raise ValueError(str(e)) from e
        if self.rank == 0:
            self._data[()].copy_(dense)
            return
        if not self._host and _kernel_dtype(self._tdtype):  # CUDA pack kernel: gather + symmetry check / symmetrization fused
            if not pack_dense_device(LAYOUT_PERMCLS, self.rank, self.dim, dense, self._buf, symmetrize):
                raise ValueError("Data array is not symmetric.")
            return
        gpu = _gpu_for_host() if _kernel_dtype(self._tdtype) else None
        if gpu is not None:  # host-resident tensor: pack on the device, keep the packed buffer on the host
            tmp = torch.zeros(self._buf.numel(), dtype=self._tdtype, device=gpu)
            if not pack_dense_device(LAYOUT_PERMCLS, self.rank, self.dim, dense.to(gpu), tmp, symmetrize):
                raise ValueError("Data array is not symmetric.")
            self._buf.copy_(tmp)
            return
        perms = list(itertools.permutations(range(self.rank)))
        for c in self._table.classes:
            if len(c) > self.dim:
                continue
            idx = self._rep_index_tensor(c)
            if symmetrize:
                acc = torch.zeros(idx.shape[0], dtype=self._tdtype, device=self.device)
                for p in perms:
                    acc += dense[tuple(idx[:, k] for k in p)]
                self._data[c].copy_(acc / len(perms))
            else:
                self._data[c].copy_(dense[tuple(idx[:, k] for k in range(self.rank))])
        if not symmetrize and not torch.allclose(self.todense(), dense, rtol=1e-5, atol=1e-8, equal_nan=True):
            raise ValueError("Data array is not symmetric.")

    # ---- public attributes ---------------------------------------------------------------------------
    @property
    def dtype(self) -> np.dtype:
        return _TORCH2NP[self._tdtype]

    @property
    def torch_dtype(self) -> torch.dtype:
        return self._tdtype

    @property
    def perm_classes(self):
        return [comb.permclass_counts_to_label(c) for c in self._table.classes]

    @property
    def size(self) -> int:
        return self.indep_size

    @property
    def packed(self) -> torch.Tensor:
        """The contiguous packed buffer (classes in order, starts padded to 32 elements)."""
        return self._buf

    @property
    def class_table(self):
        return self._table

    def keys(self):
        return self._data.keys()

    def values(self):
        return list(self._data.values())

    def items(self):
        return list(self._data.items())

    def copy(self):
        return type(self).from_packed(self.rank, self.dim, self._buf.clone())

    clone = copy

    def astype(self, dtype):
        return type(self).from_packed(self.rank, self.dim, self._buf.to(to_torch_dtype(dtype)))

    def to(self, device):
        """Move the packed buffer: 'host' (pinned) or a CUDA device."""
        if _is_host(device):
            buf = torch.empty(self._buf.shape, dtype=self._tdtype, pin_memory=torch.cuda.is_available())
            buf.copy_(self._buf)
        else:
            buf = self._buf.to(device)
        return type(self).from_packed(self.rank, self.dim, buf)

    def to_numpy_dict(self) -> Dict[Cls, np.ndarray]:
        return {c: v.detach().cpu().numpy() for c, v in self._data.items()}

    def permcls_multiplicity(self, c) -> int:
        return comb.permclass_multiplicity(_parse_class_key(c))

    # ---- iteration (symtensor/permcls_symtensor.py:934-979) ----------------------------------------------
    def indep_iter(self):
        for v in self._data.values():
            yield from v.reshape(-1)

    def indep_iter_repindex(self) -> Iterable[Tuple[int, ...]]:
        for c in self._table.classes:
            yield from self.permcls_indep_iter_repindex(c)

    def permcls_indep_iter_repindex(self, c) -> Iterable[Tuple[int, ...]]:
        c = _parse_class_key(c)
        if self.rank == 0:
            yield ()
            return
        for row in self._rep_index_tensor(c).cpu().tolist():
            yield tuple(row)

    # ---- conversion ----------------------------------------------------------------------------------
    def todense(self) -> torch.Tensor:
        """Dense ``dim**rank`` tensor on the same device (small tensors / tests only)."""
        if self.rank == 0:
            return self._data[()].clone()
        if not self._host and _kernel_dtype(self._tdtype):
            if self.rank <= 8 and self.dim ** self.rank < 2 ** 32 and self.rank * self.dim * 4 <= 40 * 1024:
                # re-order to the flat layout (N components), then the compile-time-rank unpack kernel (dim^rank elements)
                flat = torch.empty(comb.indep_size(self.rank, self.dim), dtype=self._tdtype, device=self.device)
                fn = lib.st_permcls_to_flat_f64 if self._tdtype == torch.float64 else lib.st_permcls_to_flat_f32
                with torch.cuda.device(self.device):
                    check(fn(self.rank, c_i64(self.dim), self._buf.data_ptr(), flat.data_ptr(), _stream_ptr(self.device)))
                return unpack_dense_device(1, self.rank, self.dim, flat)
            return unpack_dense_device(LAYOUT_PERMCLS, self.rank, self.dim, self._buf)
        gpu = _gpu_for_host() if _kernel_dtype(self._tdtype) else None
        if gpu is not None:  # host-resident tensor: unpack on the device
            return unpack_dense_device(LAYOUT_PERMCLS, self.rank, self.dim, self._buf.to(gpu)).cpu()
        dense = torch.zeros(self.shape, dtype=self._tdtype, device=self.device)
        perms = list(itertools.permutations(range(self.rank)))
        for c in self._table.classes:
            if len(c) > self.dim:
                continue
            idx = self._rep_index_tensor(c)
            for p in perms:
                dense[tuple(idx[:, k] for k in p)] = self._data[c]
        return dense

    # ---- indexing (symtensor/permcls_symtensor.py:724-858) -------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, str):
            return self._data[comb.permclass_label_to_counts(key)]
        if isinstance(key, (int, np.integer)):
            key = (int(key),)
        if isinstance(key, tuple):
            if any(isinstance(k, slice) for k in key):
                # only `:` slices, which just drop out of the key (symtensor/permcls_symtensor.py:735-748)
                if any(isinstance(k, slice) and k != slice(None) for k in key):
                    raise NotImplementedError("Indexing with subslicing (for example SymmetricTensor[1:3, 0,0]) is not currently implemented. "
                                              "Only slices of the type [i_1,...,i_n,:,...,:] with i_1,..., i_n all integers are allowed.")
                key = tuple(int(k) for k in key if not isinstance(k, slice))
                if len(key) == 0:
                    return self
            if len(key) < self.rank:
                # fewer indices than the rank: the rank-lowering gather (symtensor/permcls_symtensor.py:750-781) on the device
                return self.slice_fixed(key)
            c, pos = comb.convert_dense_index(self.rank, self.dim, key)
            return self._data[c][pos] if self.rank else self._data[c]
        raise KeyError(f"{key}")

    def __setitem__(self, key, value):
        if isinstance(key, slice) and key == slice(None):
            if isinstance(value, CudaPermClsSymmetricTensor):
                if value.data_alignment != self.data_alignment:
                    raise ValueError("Cannot assign to SymmetricTensor: value has an incompatible shape.")
                self._buf.copy_(value._buf)
            elif isinstance(value, SymmetricTensor):
                raise NotImplementedError("assignment of a SymmetricTensor of another format is not supported")
            else:
                self._init_from_dense(value, symmetrize=False)
            return
        if isinstance(key, str):
            c = comb.get_permclass(tuple(key))
            if c not in self._data:
                raise KeyError(f"'{key}' does not match any permutation class.\nPermutation classes: {self.perm_classes}.")
            if np.ndim(value) > 0 and len(value) != self._table.sizes[self._table.index(c)]:
                raise ValueError("Value must either be a scalar, or match the index class size.\n"
                                 f"Value size: {len(value)}\nPermutation class size: {self._table.sizes[self._table.index(c)]}")
            self._assign_class(c, value)
            return
        if isinstance(key, (int, np.integer)):
            key = (int(key),)
        c, pos = comb.convert_dense_index(self.rank, self.dim, key)
        if self.rank:
            self._data[c][pos] = value
        else:
            self._data[c].fill_(value)

    # ---- rank-0 conveniences (results of full contractions) -------------------------------------------------
    def item(self):
        if self.rank != 0:
            raise ValueError("only rank-0 tensors convert to Python scalars")
        return self._data[()].item()

    def __float__(self):
        return float(self.item())

    def __array__(self, dtype=None, copy=None):
        arr = self.todense().detach().cpu().numpy()
        return arr.astype(dtype) if dtype is not None else arr

    def __repr__(self):
        s = f"{type(self).__qualname__}(rank: {self.rank}, dim: {self.dim}, device: {'host' if self._host else self.device})"
        lines = [f"  {comb.permclass_counts_to_label(c)}: {v}" for c, v in self._data.items()]
        return "\n".join((s, *lines)) + "\n"


# names used by the reference / BASELINE.json for the torch-backed permcls class
PermClsTorchSymmetricTensor = CudaPermClsSymmetricTensor
TorchPermClsSymmetricTensor = CudaPermClsSymmetricTensor
PermClsSymmetricTensor = CudaPermClsSymmetricTensor
