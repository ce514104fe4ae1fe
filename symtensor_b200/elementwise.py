"""Elementwise ufuncs, comparisons and partial indexing on the packed device buffers, so that whole
expressions stay on the GPU:

* ``np.add(A, B)``, ``A + B``, ``2 * A``, ``np.exp(A)`` ... -- the reference's ``default_unary_ufunc`` / ``default_binary_ufunc``
  (symtensor/base.py:1146-1362) apply a ufunc class by class to ``_data``; here ONE kernel runs over the packed buffer
  (``st_elementwise_*``).  Operands: tensors of the same ``data_alignment`` and scalars (a dense array operand would need the
  dense ``d**r`` array the format exists to avoid: ``NotImplemented``, as the reference does for unsupported combinations).
* ``np.isclose / np.allclose / np.array_equal`` (symtensor/base.py:1521-1684; ``__array_function__`` protocol) --
  ``st_compare_*``.  ``isclose`` returns a tensor of the same class holding 1.0 / 0.0 (the backend stores floats only).
* ``A[i]`` / ``A[i, j, :]`` with fewer indices than the rank -- the rank-lowering gather of
  symtensor/permcls_symtensor.py:750-781 (``st_slice_*``): one thread per packed component of the result.
"""
from __future__ import annotations

from numbers import Number

import numpy as np
import torch

from ._cabi import c_i64, check, lib

_UNARY = {np.negative: 0, np.absolute: 1, np.sqrt: 2, np.square: 3, np.exp: 4, np.log: 5, np.reciprocal: 6}
_BINARY = {np.add: 0, np.subtract: 1, np.multiply: 2, np.divide: 3, np.true_divide: 3, np.maximum: 4, np.minimum: 5, np.power: 6}


def _fn(name, tdt):
    return getattr(lib, f"{name}_{'f64' if tdt == torch.float64 else 'f32'}")


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _is_scalar(x):
    return isinstance(x, Number) or (isinstance(x, (np.ndarray, np.generic, torch.Tensor)) and np.ndim(x) == 0)


class PackedElementwise:
    """Mixin of the CUDA tensor classes (``_buf``: packed device buffer, ``layout``, ``rank``, ``dim``, ``from_packed``)."""

    __array_priority__ = 1000

    # ---- ufuncs ---------------------------------------------------------------------------------------------------
    def _result_dtype(self, other) -> torch.dtype:
        if isinstance(other, PackedElementwise):
            return torch.float64 if torch.float64 in (self._buf.dtype, other._buf.dtype) else torch.float32
        if isinstance(other, (bool, int, float)):
            return self._buf.dtype  # weakly typed Python scalars
        od = np.asarray(other.detach().cpu() if isinstance(other, torch.Tensor) else other).dtype
        return torch.float64 if (od == np.float64 or self._buf.dtype == torch.float64) else torch.float32

    def _need_device(self):
        if self._buf.device.type != "cuda":
            raise RuntimeError("symtensor_b200: elementwise ops need the tensor on a CUDA device (use .to('cuda')); there is no CPU fallback")

    def _unary(self, op: int):
        self._need_device()
        out = torch.empty_like(self._buf)
        with torch.cuda.device(self._buf.device):
            check(_fn("st_elementwise_unary", self._buf.dtype)(op, self.layout, self.rank, c_i64(self.dim), self._buf.data_ptr(), out.data_ptr(),
                                                               _stream(self._buf.device)))
        return type(self).from_packed(self.rank, self.dim, out)

    def _binary(self, op: int, other, reflected: bool = False, out=None):
        self._need_device()
        tdt = self._result_dtype(other)
        a = self._buf if self._buf.dtype == tdt else self._buf.to(tdt)
        res = torch.empty_like(a) if out is None else out._buf
        with torch.cuda.device(a.device):
            if isinstance(other, PackedElementwise):
                if other.data_alignment != self.data_alignment:
                    return NotImplemented
                other._need_device()
                b = other._buf if other._buf.dtype == tdt else other._buf.to(tdt)
                x, y = (b, a) if reflected else (a, b)
                check(_fn("st_elementwise_binary", tdt)(op, 0, self.layout, self.rank, c_i64(self.dim), x.data_ptr(), y.data_ptr(), 0.0, res.data_ptr(),
                                                        _stream(a.device)))
            elif _is_scalar(other):
                s = float(other.item() if hasattr(other, "item") else other)
                check(_fn("st_elementwise_binary", tdt)(op, 2 if reflected else 1, self.layout, self.rank, c_i64(self.dim), a.data_ptr(), 0, s,
                                                        res.data_ptr(), _stream(a.device)))
            else:
                return NotImplemented
        return out if out is not None else type(self).from_packed(self.rank, self.dim, res)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        out = kwargs.pop("out", None)
        if isinstance(out, tuple):
            out = out[0] if len(out) == 1 else NotImplemented
        if method != "__call__" or kwargs or out is NotImplemented:
            return NotImplemented
        if ufunc in _UNARY and len(inputs) == 1:
            res = self._unary(_UNARY[ufunc])
            if out is not None:
                out._buf.copy_(res._buf)
                return out
            return res
        if ufunc in _BINARY and len(inputs) == 2:
            a, b = inputs
            if a is self:
                return self._binary(_BINARY[ufunc], b, out=out if isinstance(out, PackedElementwise) else None)
            return self._binary(_BINARY[ufunc], a, reflected=True, out=out if isinstance(out, PackedElementwise) else None)
        return NotImplemented

    def __neg__(self):
        return self._unary(0)

    def __abs__(self):
        return self._unary(1)

    def __add__(self, o):
        return self._binary(0, o)

    def __radd__(self, o):
        return self._binary(0, o, reflected=True)

    def __sub__(self, o):
        return self._binary(1, o)

    def __rsub__(self, o):
        return self._binary(1, o, reflected=True)

    def __mul__(self, o):
        return self._binary(2, o)

    def __rmul__(self, o):
        return self._binary(2, o, reflected=True)

    def __truediv__(self, o):
        return self._binary(3, o)

    def __rtruediv__(self, o):
        return self._binary(3, o, reflected=True)

    def __pow__(self, o):
        if isinstance(o, (int, float)) and o == 2:
            return self._unary(3)  # like NumPy, x ** 2 is the exact square
        return self._binary(6, o)

    def __iadd__(self, o):
        r = self._binary(0, o, out=self if self._result_dtype(o) == self._buf.dtype else None)
        if r is NotImplemented:
            return r
        if r is not self:
            self._buf.copy_(r._buf)
        return self

    def __isub__(self, o):
        r = self._binary(1, o, out=self if self._result_dtype(o) == self._buf.dtype else None)
        if r is NotImplemented:
            return r
        if r is not self:
            self._buf.copy_(r._buf)
        return self

    def __imul__(self, o):
        r = self._binary(2, o, out=self if self._result_dtype(o) == self._buf.dtype else None)
        if r is NotImplemented:
            return r
        if r is not self:
            self._buf.copy_(r._buf)
        return self

    # ---- comparisons (NEP 18) --------------------------------------------------------------------------------------
    def _compare(self, other, mode: int, rtol=1e-5, atol=1e-8, equal_nan=False, want_mask=False):
        self._need_device()
        dev = self._buf.device
        if isinstance(other, PackedElementwise):
            if other.data_alignment != self.data_alignment:
                return None
            tdt = self._result_dtype(other)
            a = self._buf if self._buf.dtype == tdt else self._buf.to(tdt)
            b = other._buf.to(device=dev, dtype=tdt)
            scalar, s = 0, 0.0
        elif _is_scalar(other):
            tdt = self._buf.dtype
            a, b, scalar, s = self._buf, None, 1, float(other.item() if hasattr(other, "item") else other)
        else:
            return None
        flag = torch.ones(1, dtype=torch.int32, device=dev)
        mask = torch.empty_like(a) if want_mask else None
        with torch.cuda.device(dev):
            check(_fn("st_compare", tdt)(mode, self.layout, self.rank, c_i64(self.dim), a.data_ptr(), b.data_ptr() if b is not None else 0, scalar, s,
                                         float(rtol), float(atol), int(bool(equal_nan)), mask.data_ptr() if mask is not None else 0, flag.data_ptr(),
                                         _stream(dev)))
        if want_mask:
            return type(self).from_packed(self.rank, self.dim, mask)
        return bool(flag.item())

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.isclose, np.allclose, np.array_equal):
            a, b = args[0], args[1]
            extra = dict(zip(("rtol", "atol", "equal_nan"), args[2:]))
            extra.update(kwargs)
            me, other = (a, b) if isinstance(a, PackedElementwise) else (b, a)
            if func is np.array_equal:
                if isinstance(other, PackedElementwise) and (other.rank != me.rank or other.dim != me.dim):
                    return False
                res = me._compare(other, 0)
            elif func is np.allclose:
                if isinstance(other, PackedElementwise) and (other.rank != me.rank or other.dim != me.dim):
                    return False
                res = me._compare(other, 1, **extra)
            else:
                res = me._compare(other, 1, want_mask=True, **extra)
            return NotImplemented if res is None else res
        return NotImplemented

    # ---- partial indexing ------------------------------------------------------------------------------------------
    def slice_fixed(self, fixed):
        """``A[i_1, ..., i_n]`` with n < rank: the rank-(rank - n) tensor ``B[K] = A[K + fixed]``."""
        self._need_device()
        fixed = [int(i) for i in fixed]
        if any(i < 0 or i >= self.dim for i in fixed):
            raise IndexError(f"index {fixed} out of range for dimension {self.dim}")
        n = len(fixed)
        if not 0 < n < self.rank:
            raise ValueError("slice_fixed needs between 1 and rank - 1 indices")
        dev = self._buf.device
        new_rank = self.rank - n
        proto = type(self).from_packed.__func__
        from . import combinatorics as comb
        total = comb.class_table(new_rank, self.dim).total if self.layout == 0 else comb.indep_size(new_rank, self.dim)
        out = torch.empty(total, dtype=self._buf.dtype, device=dev)
        fx = torch.tensor(fixed, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(_fn("st_slice", self._buf.dtype)(self.layout, self.rank, c_i64(self.dim), n, fx.data_ptr(), self._buf.data_ptr(), out.data_ptr(),
                                                   _stream(dev)))
        return proto(type(self), new_rank, self.dim, out)
