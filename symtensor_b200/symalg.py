"""The symmetric-algebra dispatch surface: same names, argument meaning and errors as ``symtensor.symalg``.

    contract_all_indices_with_vector(A, x)      symtensor/symalg.py:505-527
    contract_all_indices_with_matrix(A, W)      symtensor/symalg.py:475-496
    tensordot(a, b, axes=2)                     symtensor/symalg.py:427-459   (symmetrized)
    multiply.outer(a, b) / add / subtract       symtensor/symalg.py:101-171, 193-195, 294-316

Each call is routed to the implementation registered by the argument classes (subclasses first, then left to
right -- ``ufunc_dispatch``'s rule, symalg.py:120-171).  Unlike the reference there is NO dense fallback: the
reference's defaults densify to ``d**r`` and average ``r!`` transposes on the CPU; here an op without a
registered CUDA implementation raises ``TypeError``.
"""
from __future__ import annotations

import numpy as np

from .base import SymmetricTensor, result_array  # noqa: F401  (re-exported)


def _ordered_providers(args):
    """Distinct SymmetricTensor types among ``args``, subclasses ahead of their superclasses."""
    out = []
    for a in args:
        t = type(a)
        if not isinstance(a, SymmetricTensor) or t in out:
            continue
        for i, u in enumerate(out):
            if t is not u and issubclass(t, u):
                out.insert(i, t)
                break
        else:
            out.append(t)
    return out


class _ArrayFunction:
    """A dispatched function object; it is also the registry key (like NumPy's public wrappers)."""

    def __init__(self, name, dispatcher, doc):
        self.__name__ = name
        self.__qualname__ = name
        self.__doc__ = doc
        self._dispatcher = dispatcher

    def __call__(self, *args, **kwargs):
        relevant = self._dispatcher(*args, **kwargs)
        for t in _ordered_providers(relevant):
            impl = t._HANDLED_FUNCTIONS.get(self)
            if impl is None:
                continue
            res = impl(*args, **kwargs)
            if res is not NotImplemented:
                return res
        raise TypeError(f"no implementation of symalg.{self.__name__} for argument types "
                        f"{[type(a).__name__ for a in relevant]} (symtensor_b200 has no dense CPU fallback)")

    def __repr__(self):
        return f"<symalg function {self.__name__}>"


contract_all_indices_with_vector = _ArrayFunction(
    "contract_all_indices_with_vector", lambda symtensor, x: (symtensor, x),
    "sum_{i1..ir} A[i1..ir] x[i1]...x[ir]; returns a rank-0 tensor of A's class (int 0 for an all-zero x).")

contract_all_indices_with_matrix = _ArrayFunction(
    "contract_all_indices_with_matrix", lambda symtensor, W: (symtensor, W),
    "C[j1..jr] = sum A[i1..ir] W[i1,j1]...W[ir,jr]; returns a tensor of A's class.")

tensordot = _ArrayFunction(
    "tensordot", lambda a, b, axes=2: (a, b),
    "Symmetrized tensordot: Sym(sum over `axes` contracted index pairs).")


class UfuncWrapper:
    """``symalg.multiply`` etc.: calling it is the plain NumPy ufunc, ``.outer`` is the symmetrized outer
    product dispatched through the classes' ``_HANDLED_UFUNCS['outer']`` registries."""

    def __init__(self, ufunc):
        self.ufunc = ufunc
        self.__name__ = ufunc.__name__
        self.signature = ufunc.signature

    def __call__(self, *args, **kwargs):
        return self.ufunc(*args, **kwargs)

    def outer(self, a, b, **kwargs):
        for t in _ordered_providers((a, b)):
            impl = t._HANDLED_UFUNCS["outer"].get(self)
            if impl is None:
                continue
            res = impl(self, a, b, **kwargs)
            if res is not NotImplemented:
                return res
        raise TypeError(f"no implementation of symalg.{self.__name__}.outer for argument types "
                        f"({type(a).__name__}, {type(b).__name__}) (symtensor_b200 has no dense CPU fallback)")

    def __repr__(self):
        return f"<symalg ufunc {self.__name__}>"


add = UfuncWrapper(np.add)
subtract = UfuncWrapper(np.subtract)
multiply = UfuncWrapper(np.multiply)


def contract_tensor_list(symtensor, tensor_list, n_times: int = 1, rule: str = "second_half"):
    """symtensor/symalg.py:556-642: for A = ``symtensor`` and chi_i = ``tensor_list[i]`` (all symmetric, same shape)

        B = Sym[ sum over the last n indices  A[..., i_1, ..., i_n] (x) chi_{i_1} (x) ... (x) chi_{i_n} ]

    computed as the reference does -- ``C += reduce(multiply.outer, (chi_i for i in idx), A[idx])`` over the index tuples -- but
    with every piece on the device: ``A[idx]`` is the rank-lowering gather (``st_slice_*``), the products are the symmetrized
    outer kernels, the accumulation an elementwise add on the packed buffer.  ``rule='second_half'`` (the reference's default)
    sums only over indices ``>= ceil(dim / 2)``; the reference raises ``NameError`` there because ``math`` is not imported
    (symalg.py:628) -- this implementation computes what that branch was written to compute.  ``rule='all'`` (any other value in
    the reference) sums over all indices."""
    import math
    from functools import reduce
    from itertools import product

    tensor_list = list(tensor_list)
    if not isinstance(symtensor, SymmetricTensor) or not all(isinstance(x, SymmetricTensor) for x in tensor_list):
        return NotImplemented
    cls = result_array(symtensor, *tensor_list)
    A = symtensor
    if n_times > A.rank:
        raise ValueError(f"n_times is {n_times}, but cannot do more contractions than {A.rank} with tensor of rank {A.rank}")
    if len(tensor_list) != A.dim:
        raise ValueError("`tensor_list` emulates the first dimension of a tensor, and therefore its length must match the dimenion of "
                         f"`symtensor`.\nSymtensor dim     : {A.dim}\nLength tensor list: {len(tensor_list)}")
    ranks, dims = {x.rank for x in tensor_list}, {x.dim for x in tensor_list}
    if len(ranks) > 1 or len(dims) > 1:
        raise ValueError(f"Tensors in `tensor_list` do not all have the same shape:\n{[x.shape for x in tensor_list]}")
    chi_rank, chi_dim = next(iter(ranks)), next(iter(dims))
    if chi_dim != A.dim:
        raise ValueError("Tensors in `tensor_list` do not have the same dimension as `symtensor`.")
    if A.rank == 1 and n_times == 1:
        acc = None
        for i in range(A.dim):
            term = tensor_list[i] * A[i]
            acc = term if acc is None else acc + term
        return acc if acc is not None else cls(tensor_list[0].rank, tensor_list[0].dim)
    if rule == "second_half":
        indices = product(range(math.ceil(A.dim / 2), A.dim), repeat=n_times)
    else:
        indices = product(range(A.dim), repeat=n_times)
    C = None
    for idx in indices:
        head = A[idx] if n_times < A.rank else A[idx]  # rank-lowering gather (a scalar tensor entry when n_times == rank)
        term = reduce(multiply.outer, (tensor_list[i] for i in idx), head)
        C = term if C is None else C + term
    if C is None:  # no index to sum over (dim 0 or 1 with 'second_half'): a zero tensor of the result shape
        C = cls(rank=A.rank + n_times * (chi_rank - 1), dim=A.dim, device=getattr(A, "device", None))
    return C
