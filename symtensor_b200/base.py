"""Dispatch core of the standalone package: a small mirror of the reference's registry mechanism.

In the reference, ops are found through class-level registries filled by decorators
(``SymmetricTensor.implements`` symtensor/base.py:1057-1063, ``implements_ufunc.outer`` :259-322; every
subclass gets child maps over its MRO, :682-698).  This module keeps the same surface -- ``@Cls.implements(f)``,
``@Cls.implements_ufunc.outer(symalg.multiply)`` -- without NumPy's private override machinery, so that the
GPU backend classes are written exactly like a reference backend mixin.  ``plugin.py`` registers the same
implementations into the real ``symtensor`` package when it is importable.
"""
from __future__ import annotations

from collections import ChainMap
from typing import Callable, Tuple

import numpy as np

from . import combinatorics as comb


class _UfuncRegistry:
    """``cls.implements_ufunc.outer(*ufuncs)`` -> decorator storing into ``cls._HANDLED_UFUNCS['outer']``."""

    def __init__(self, owner):
        self._owner = owner

    def _register(self, method: str, ufuncs):
        def decorator(f):
            for uf in ufuncs:
                self._owner._HANDLED_UFUNCS[method][uf] = f
            return f
        return decorator

    def outer(self, *ufuncs):
        return self._register("outer", ufuncs)

    def __call__(self, *ufuncs):
        return self._register("__call__", ufuncs)


class _implements_ufunc_descriptor:
    def __get__(self, obj, owner):
        return _UfuncRegistry(owner)


class SymmetricTensor:
    """Abstract base: rank / dim / dtype bookkeeping and the op registries."""

    data_format = "None"
    _HANDLED_FUNCTIONS = ChainMap()
    _HANDLED_UFUNCS = {"outer": ChainMap(), "__call__": ChainMap()}
    implements_ufunc = _implements_ufunc_descriptor()

    rank: int
    dim: int

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        # fresh child maps: a subclass registration shadows its parents' without touching them
        parents_f = [b._HANDLED_FUNCTIONS for b in cls.__mro__[1:] if "_HANDLED_FUNCTIONS" in vars(b)]
        cls._HANDLED_FUNCTIONS = ChainMap({}, *[m for p in parents_f for m in p.maps])
        cls._HANDLED_UFUNCS = {}
        for method in ("outer", "__call__"):
            parents_u = [b._HANDLED_UFUNCS[method] for b in cls.__mro__[1:] if "_HANDLED_UFUNCS" in vars(b)]
            cls._HANDLED_UFUNCS[method] = ChainMap({}, *[m for p in parents_u for m in p.maps])

    @classmethod
    def implements(cls, function) -> Callable:
        """Register an implementation of a ``symalg`` function for this class (and its subclasses)."""
        def decorator(f):
            cls._HANDLED_FUNCTIONS[function] = f
            return f
        return decorator

    # ---- shape bookkeeping (symtensor/base.py:805-844)
    @property
    def ndim(self) -> int:
        return self.rank

    @property
    def shape(self) -> Tuple[int, ...]:
        return (self.dim,) * self.rank

    @property
    def dense_size(self) -> int:
        return self.dim ** self.rank

    @property
    def indep_size(self) -> int:
        return comb.indep_size(self.rank, self.dim)

    @property
    def data_alignment(self):
        return (self.data_format, self.rank, self.dim)

    def transpose(self, *axes):
        return self

    @property
    def T(self):
        return self


def common_superclass(*types):
    """Most specific class all ``types`` share (symtensor/utils.py:63-72)."""
    if not types:
        return np.ndarray
    mros = [t.__mro__ for t in types]
    for c in mros[0]:
        if all(c in m for m in mros[1:]):
            return c
    return object


def result_array(*arrays_and_types):
    """Class promotion (symtensor/base.py:1757-1794): ndarray unless a SymmetricTensor takes part, then the
    most specific common SymmetricTensor subclass."""
    types = [a if isinstance(a, type) else type(a) for a in arrays_and_types]
    sym = tuple(t for t in types if issubclass(t, SymmetricTensor))
    return common_superclass(*sym) if sym else np.ndarray
