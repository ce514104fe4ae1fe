"""Dense oracle: the reference's ACTUAL algorithm for the four ops, restated in NumPy.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  For PermCls / Flat tensors the reference has no packed
implementation: every op converts to a dense ``d**r`` array, calls a NumPy routine, averages over all ``r!``
axis permutations and re-packs (``symtensor/symalg.py:206-283, 294-316, 427-459, 475-496, 505-527``).  This
module follows that recipe step by step (vectorised where the reference uses Python loops, which changes
no value) so it can only be used for small ``d**r``.  A tensor is the reference's ``_data`` mapping
``{class tuple: 1-D array | 0-d scalar}`` in ``perm_classes`` order.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, Tuple

import numpy as np

from . import index_oracle as io

Packed = Dict[Tuple[int, ...], np.ndarray]


def class_array(v, cls, dim, dtype=None) -> np.ndarray:
    """Expand a class entry to a full 1-D array: 0-d scalars broadcast (``permcls_symtensor.py:943-949``);
    a missing / empty class of a dict-constructed tensor is treated as zeros."""
    size = io.permclass_size(cls, dim)
    v = np.asarray(v)
    if v.ndim == 0:
        return np.full(size, v, dtype=dtype or v.dtype)
    if v.shape == (0,) and size:
        return np.zeros(size, dtype=dtype or v.dtype)
    assert v.shape == (size,), (cls, v.shape, size)
    return v.astype(dtype) if dtype is not None else v


def todense(data: Packed, rank: int, dim: int, dtype=np.float64) -> np.ndarray:
    """Scatter every stored component to all of its index permutations
    (``permcls_symtensor.py:883-887`` with ``base.py:919-935`` / ``utils.py:647-650``)."""
    if rank == 0:
        return np.asarray(data[()], dtype=dtype).reshape(())
    dense = np.zeros((dim,) * rank, dtype=dtype)
    perms = list(itertools.permutations(range(rank)))
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue
        vals = class_array(data.get(cls, np.zeros(0)), cls, dim, dtype)
        rep = io.class_repindex(cls, dim)  # [size, rank]
        for p in perms:
            dense[tuple(rep[:, k] for k in p)] = vals
    return dense


def symmetrize(t: np.ndarray) -> np.ndarray:
    """Plain average over all ``ndim!`` transposes (``symtensor/utils.py:507-532``)."""
    n = t.ndim
    if n <= 1:
        return t
    acc = np.zeros_like(t)
    for p in itertools.permutations(range(n)):
        acc = acc + t.transpose(p)
    return acc / math.factorial(n)


def repack(dense: np.ndarray, rank: int, dim: int) -> Packed:
    """Gather the representative of every stored component (``permcls_symtensor.py:599-618``)."""
    if rank == 0:
        return {(): np.asarray(dense).reshape(())}
    out: Packed = {}
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue  # the reference leaves these as empty arrays (``permcls_symtensor.py:618, 666-667``)
        rep = io.class_repindex(cls, dim)
        out[cls] = dense[tuple(rep[:, k] for k in range(rank))]
    return out


def tensordot(a: Packed, ra: int, b, rb: int, dim: int, axes=2, dtype=np.float64):
    """Symmetrized tensordot (``symtensor/symalg.py:427-459``).  ``b`` may be packed (dict) or a dense
    ndarray (e.g. a vector).  Returns (packed result, rank)."""
    da = todense(a, ra, dim, dtype) if isinstance(a, dict) else np.asarray(a, dtype=dtype)
    db = todense(b, rb, dim, dtype) if isinstance(b, dict) else np.asarray(b, dtype=dtype)
    out = symmetrize(np.tensordot(da, db, axes))
    r = out.ndim
    return repack(out, r, dim if r else 1), r


def contract_all_indices_with_vector(a: Packed, rank: int, dim: int, x, dtype=np.float64) -> float:
    """r-fold symmetrized ``tensordot(., x, axes=1)`` (``symtensor/symalg.py:505-527``), including the
    re-pack / densify round trip between the steps."""
    x = np.asarray(x, dtype=dtype)
    cur, r = a, rank
    for _ in range(rank):
        cur, r = tensordot(cur, r, x, 1, dim, axes=1, dtype=dtype)
    return float(np.asarray(cur[()]))


def contract_all_indices_with_matrix(a: Packed, rank: int, dim: int, W, dtype=np.float64) -> Packed:
    """C[j1..jr] = sum A[i1..ir] prod_k W[i_k, j_k]  (``symtensor/symalg.py:475-496``).

    The reference evaluates one ``np.einsum`` over all 2r letters; here the same sum is taken mode by
    mode (identical up to floating-point summation order)."""
    t = todense(a, rank, dim, dtype)
    W = np.asarray(W, dtype=dtype)
    for _ in range(rank):
        # contract the current first axis with W's first axis; the new axis is appended last
        t = np.tensordot(t, W, axes=([0], [0]))
    return repack(t, rank, dim)


def outer(a: Packed, ra: int, b: Packed, rb: int, dim: int, dtype=np.float64) -> Packed:
    """Symmetrized outer product (``symtensor/symalg.py:294-316`` via ``symmetrized_op`` ``:206-283``)."""
    da, db = todense(a, ra, dim, dtype), todense(b, rb, dim, dtype)
    return repack(symmetrize(np.multiply.outer(da, db)), ra + rb, dim)
