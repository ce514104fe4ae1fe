"""Packed oracle: the four ops computed directly on the packed components (NumPy, fp64).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  These are the packed-space formulas that are
mathematically identical to the reference's dense defaults (op semantics
``symtensor/symalg.py:206-283, 294-316, 427-459, 475-496, 505-527``).  They exist because the reference's
own path needs a dense ``d**r`` array and cannot run the larger configurations.  Validated against
``dense_oracle`` and against the unmodified reference (``tests/golden``) in ``tests/test_oracle_ops.py``.

A tensor is the reference's ``_data`` mapping ``{class tuple: 1-D array | 0-d scalar}``.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, Tuple

import numpy as np

from . import index_oracle as io
from .dense_oracle import class_array

Packed = Dict[Tuple[int, ...], np.ndarray]


def _binom_table(n: int, k: int) -> np.ndarray:
    t = np.zeros((n + 1, k + 1), dtype=np.int64)
    for i in range(n + 1):
        for j in range(min(i, k) + 1):
            t[i, j] = math.comb(i, j)
    return t


def flat_rank_array(idx: np.ndarray, dim: int) -> np.ndarray:
    """Vectorised ``flat_symtensor.index_of_multicombination`` (``flat_symtensor.py:39-50``) for an array
    [n, r] of SORTED multi-indices."""
    n, r = idx.shape
    if r == 0:
        return np.zeros(n, dtype=np.int64)
    bt = _binom_table(dim + r, r)
    pos = np.full(n, math.comb(dim + r - 1, r) - 1, dtype=np.int64)
    for k in range(r):
        ck = idx[:, r - 1 - k].astype(np.int64)
        pos -= bt[dim - 1 + k - ck, k + 1]
    return pos


def permcls_to_flat(data: Packed, rank: int, dim: int, dtype=np.float64) -> np.ndarray:
    """Re-order a permcls tensor into flat (``combinations_with_replacement``) order."""
    out = np.zeros(io.indep_size(rank, dim), dtype=dtype)
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue
        rep = np.sort(io.class_repindex(cls, dim), axis=1)
        out[flat_rank_array(rep, dim)] = class_array(data.get(cls, np.zeros(0)), cls, dim, dtype)
    return out


def flat_to_permcls(flat: np.ndarray, rank: int, dim: int) -> Packed:
    out: Packed = {}
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue
        rep = np.sort(io.class_repindex(cls, dim), axis=1)
        out[cls] = flat[flat_rank_array(rep, dim)]
    return out


# --------------------------------------------------------------------------------------------------
def contract_all_indices_with_vector(data: Packed, rank: int, dim: int, x, dtype=np.float64) -> float:
    """s = sum_cls gamma_cls * sum_p A_cls[p] * prod_j x[v_j(p)]**m_j   (SURVEY.md A.3)."""
    x = np.asarray(x, dtype=dtype)
    if rank == 0:
        return float(np.asarray(data[()]))
    total = 0.0
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue
        vals = class_array(data.get(cls, np.zeros(0)), cls, dim, dtype)
        v = io.class_values(cls, dim)
        w = np.ones(v.shape[0], dtype=dtype)
        for j, m in enumerate(cls):
            w = w * x[v[:, j]] ** m
        total += io.permclass_multiplicity(cls) * math.fsum(vals * w)
    return float(total)


def contract_vec_flat(flat: np.ndarray, rank: int, dim: int, x) -> float:
    """Same sum over the flat layout: s = sum_p mult(I_p) * A[p] * prod_k x[I_p[k]]
    (layout ``flat_symtensor.py:219-220``, multiplicity ``:59-74``)."""
    x = np.asarray(x, dtype=np.float64)
    idx = io.flat_indices(rank, dim)
    w = np.prod(x[idx], axis=1) if rank else np.ones(1)
    mult = np.array([io.flat_multiplicity(t) for t in idx.tolist()], dtype=np.float64)
    return float(math.fsum(np.asarray(flat, dtype=np.float64) * w * mult))


def outer(a: Packed, ra: int, b: Packed, rb: int, dim: int, dtype=np.float64) -> Packed:
    """C_K = C(n,ra)^-1 * sum over all position subsets S (|S| = ra) of A[K_S] * B[K_S^c]."""
    n = ra + rb
    af, bf = permcls_to_flat(a, ra, dim, dtype), permcls_to_flat(b, rb, dim, dtype)
    subsets = list(itertools.combinations(range(n), ra))
    out: Packed = {}
    for cls in io.perm_classes(n):
        if len(cls) > dim:
            continue
        K = np.sort(io.class_repindex(cls, dim), axis=1)
        acc = np.zeros(K.shape[0], dtype=dtype)
        for S in subsets:
            Sc = [p for p in range(n) if p not in S]
            acc += af[flat_rank_array(K[:, list(S)], dim)] * bf[flat_rank_array(K[:, Sc], dim)]
        out[cls] = acc / len(subsets)
    return out


def tensordot(a: Packed, ra: int, b: Packed, rb: int, dim: int, k: int, dtype=np.float64):
    """C_K = C(n, ra-k)^-1 sum_S sum_{J in [d]^k} A[K_S, J] B[J, K_S^c],  n = ra + rb - 2k.
    Returns (packed, n); for n == 0 the packed result is ``{(): 0-d}`` (dim 1 in the reference)."""
    n = ra + rb - 2 * k
    na = ra - k
    af, bf = permcls_to_flat(a, ra, dim, dtype), permcls_to_flat(b, rb, dim, dtype)
    J = io.flat_indices(k, dim)  # sorted contracted tuples ...
    Jmult = np.array([io.flat_multiplicity(t) for t in J.tolist()], dtype=dtype)  # ... and their counts
    subsets = list(itertools.combinations(range(n), na))
    out: Packed = {}
    for cls in (io.perm_classes(n) if n else [()]):
        if len(cls) > dim:
            continue
        K = np.sort(io.class_repindex(cls, dim), axis=1) if n else np.zeros((1, 0), dtype=np.int32)
        acc = np.zeros(K.shape[0], dtype=dtype)
        for S in subsets:
            Sc = [p for p in range(n) if p not in S]
            KS, KSc = K[:, list(S)], K[:, Sc]
            for j, jm in zip(J, Jmult):
                jj = np.broadcast_to(j, (K.shape[0], k))
                ia = np.sort(np.concatenate([KS, jj], axis=1), axis=1)
                ib = np.sort(np.concatenate([jj, KSc], axis=1), axis=1)
                acc += jm * af[flat_rank_array(ia, dim)] * bf[flat_rank_array(ib, dim)]
        res = acc / len(subsets)
        out[cls] = res if n else res.reshape(())
    return out, n


def _dense_from_flat(flat: np.ndarray, rank: int, dim: int) -> np.ndarray:
    grid = np.indices((dim,) * rank).reshape(rank, -1).T  # all d**r multi-indices
    return flat[flat_rank_array(np.sort(grid, axis=1), dim)].reshape((dim,) * rank)


def contract_all_indices_with_matrix(a: Packed, rank: int, dim: int, W, dtype=np.float64) -> Packed:
    """C = (W^T)^{(x) r} . A via r mode products (W contracted on its FIRST axis), gathered back to the
    packed representatives.  Unlike ``dense_oracle`` no r!-symmetrization / Python scatter is involved:
    the dense array is a gather from the flat order."""
    W = np.asarray(W, dtype=dtype)
    t = _dense_from_flat(permcls_to_flat(a, rank, dim, dtype), rank, dim)
    for _ in range(rank):
        t = np.tensordot(t, W, axes=([0], [0]))
    out: Packed = {}
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue
        rep = io.class_repindex(cls, dim)
        out[cls] = t[tuple(rep[:, q] for q in range(rank))]
    return out


def semi_packed_chain_flops(rank: int, dim: int) -> int:
    """Algorithmic flops of the partially-symmetric mode chain."""
    return sum(2 * dim * math.comb(dim + k - 1, k) * dim * math.comb(dim + rank - k - 2, rank - k - 1)
               for k in range(rank))
