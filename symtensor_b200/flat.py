"""``CudaFlatSymmetricTensor``: the reference's flat packed format (symtensor/flat_symtensor.py:77-274) in HBM.

One 1-D tensor of length C(d+r-1, r) whose position p holds the component with the p-th sorted multi-index of
``itertools.combinations_with_replacement(range(dim), rank)`` (flat_symtensor.py:39-50, 219-220).
"""
from __future__ import annotations

import itertools
from numbers import Number
from typing import Iterable, Tuple

import numpy as np
import torch

from . import combinatorics as comb
from ._cabi import LAYOUT_FLAT, c_i64, check, lib
from .base import SymmetricTensor
from .elementwise import PackedElementwise
from .permcls import _TORCH2NP, _gpu_for_host, _is_host, _kernel_dtype, _stream_ptr, pack_dense_device, to_torch_dtype, unpack_dense_device


class CudaFlatSymmetricTensor(PackedElementwise, SymmetricTensor):
    data_format = "Flat"
    layout = LAYOUT_FLAT
    array_type = torch.Tensor

    def __init__(self, rank: int, dim: int, data=None, *, dtype=None, symmetrize: bool = False, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "host"
        self._host = _is_host(device)
        self.device = torch.device("cpu") if self._host else torch.device(device)
        self.rank, self.dim = int(rank), int(dim)
        n = self.size
        if data is None:
            data = np.float64(0)
        if isinstance(data, Number):
            self._tdtype = to_torch_dtype(dtype if dtype is not None else np.asarray(data).dtype)
            self._buf = self._empty(n).fill_(data)
            return
        t = data if isinstance(data, torch.Tensor) else torch.as_tensor(np.asarray(data))
        self._tdtype = to_torch_dtype(dtype if dtype is not None else t.dtype)
        if tuple(t.shape) == (n,):
            self._buf = self._empty(n)
            self._buf.copy_(t)
            return
        if tuple(t.shape) != self.shape:
            raise RuntimeError(f"data must be scalar or array of shape {(n,)} or {self.shape}")
        dense = t.to(self.device, self._tdtype)
        self._buf = self._empty(n)
        if self.rank and not self._host and _kernel_dtype(self._tdtype):  # CUDA pack kernel
            if not pack_dense_device(LAYOUT_FLAT, self.rank, self.dim, dense, self._buf, symmetrize):
                raise RuntimeError("data is not symmetric")
            return
        gpu = _gpu_for_host() if self.rank and _kernel_dtype(self._tdtype) else None
        if gpu is not None:  # host-resident tensor: pack on the device, keep the packed buffer on the host
            tmp = torch.zeros(n, dtype=self._tdtype, device=gpu)
            if not pack_dense_device(LAYOUT_FLAT, self.rank, self.dim, dense.to(gpu), tmp, symmetrize):
                raise RuntimeError("data is not symmetric")
            self._buf.copy_(tmp)
            return
        idx = self._rep_index_tensor()
        if symmetrize:
            perms = list(itertools.permutations(range(self.rank)))
            acc = torch.zeros(n, dtype=self._tdtype, device=self.device)
            for p in perms:
                acc += dense[tuple(idx[:, k] for k in p)]
            self._buf.copy_(acc / len(perms))
        else:
            self._buf.copy_(dense[tuple(idx[:, k] for k in range(self.rank))] if self.rank else dense.reshape(1))
            if self.rank and not torch.allclose(self.todense(), dense, rtol=1e-5, atol=1e-8, equal_nan=True):
                raise RuntimeError("data is not symmetric")

    def _empty(self, n):
        if self._host:
            return torch.zeros(n, dtype=self._tdtype, pin_memory=torch.cuda.is_available())
        return torch.zeros(n, dtype=self._tdtype, device=self.device)

    @classmethod
    def from_packed(cls, rank: int, dim: int, buf: torch.Tensor):
        self = cls.__new__(cls)
        self.rank, self.dim = int(rank), int(dim)
        self._tdtype = to_torch_dtype(buf.dtype)
        self._host = buf.device.type == "cpu"
        self.device = buf.device
        assert tuple(buf.shape) == (self.size,)
        self._buf = buf
        return self

    def _rep_index_tensor(self) -> torch.Tensor:
        n = self.size
        if self._host or self.rank == 0:
            idx = [comb.flat_unrank(self.rank, self.dim, p) for p in range(n)]
            return torch.tensor(idx, dtype=torch.int64).reshape(n, self.rank).to(self.device)
        out = torch.empty((n, self.rank), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.st_flat_unrank(self.rank, c_i64(self.dim), c_i64(0), c_i64(n), out.data_ptr(), _stream_ptr(self.device)))
        return out.to(torch.int64)

    @property
    def size(self) -> int:
        return comb.indep_size(self.rank, self.dim)

    @property
    def dtype(self) -> np.dtype:
        return _TORCH2NP[self._tdtype]

    @property
    def torch_dtype(self):
        return self._tdtype

    @property
    def packed(self) -> torch.Tensor:
        return self._buf

    @property
    def _data(self):
        return self._buf

    def indep_iter(self):
        return iter(self._buf)

    def indep_iter_repindex(self) -> Iterable[Tuple[int, ...]]:
        for row in self._rep_index_tensor().cpu().tolist():
            yield tuple(row)

    def copy(self):
        return type(self).from_packed(self.rank, self.dim, self._buf.clone())

    def to(self, device):
        if _is_host(device):
            buf = torch.empty(self._buf.shape, dtype=self._tdtype, pin_memory=torch.cuda.is_available())
            buf.copy_(self._buf)
        else:
            buf = self._buf.to(device)
        return type(self).from_packed(self.rank, self.dim, buf)

    def todense(self) -> torch.Tensor:
        if self.rank == 0:
            return self._buf.reshape(()).clone()
        if not self._host and _kernel_dtype(self._tdtype):
            return unpack_dense_device(LAYOUT_FLAT, self.rank, self.dim, self._buf)
        gpu = _gpu_for_host() if _kernel_dtype(self._tdtype) else None
        if gpu is not None:  # host-resident tensor: unpack on the device
            return unpack_dense_device(LAYOUT_FLAT, self.rank, self.dim, self._buf.to(gpu)).cpu()
        dense = torch.zeros(self.shape, dtype=self._tdtype, device=self.device)
        idx = self._rep_index_tensor()
        for p in itertools.permutations(range(self.rank)):
            dense[tuple(idx[:, k] for k in p)] = self._buf
        return dense

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            key = (int(key),)
        if isinstance(key, tuple) and any(isinstance(k, slice) for k in key):
            if any(isinstance(k, slice) and k != slice(None) for k in key):
                raise NotImplementedError("only `:` slices are supported ([i_1, ..., i_n, :, ..., :])")
            key = tuple(int(k) for k in key if not isinstance(k, slice))
            if len(key) == 0:
                return self
        if not isinstance(key, tuple) or len(key) > self.rank:
            raise KeyError(f"{key}")
        if len(key) < self.rank:
            return self.slice_fixed(key)  # rank-lowering gather on the device
        return self._buf[comb.flat_rank(self.dim, key)]

    def __setitem__(self, key, value):
        if isinstance(key, (int, np.integer)):
            key = (int(key),)
        self._buf[comb.flat_rank(self.dim, key)] = value

    def item(self):
        if self.rank != 0:
            raise ValueError("only rank-0 tensors convert to Python scalars")
        return self._buf[0].item()

    def __float__(self):
        return float(self.item())

    def __array__(self, dtype=None, copy=None):
        arr = self.todense().detach().cpu().numpy()
        return arr.astype(dtype) if dtype is not None else arr

    def __repr__(self):
        return f"{type(self).__qualname__}(rank: {self.rank}, dim: {self.dim})\n  {self._buf}\n"


FlatSymmetricTensor = CudaFlatSymmetricTensor
