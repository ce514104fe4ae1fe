"""Integer oracle: permutation classes, their sizes / multiplicities and the storage order.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Pure Python / NumPy restatement of the reference's
index bookkeeping; every function cites the reference code it follows (paths relative to
``/root/reference``).  Two independent statements of the storage order are kept on purpose:

* ``sigma_index_iter``  -- the reference's generator semantics, position by position
  (``symtensor/permcls_symtensor.py:288-347``);
* ``permcls_rank`` / ``permcls_unrank`` / ``class_values`` -- the closed form that the
  CUDA enumerator implements.

``tests/test_oracle_index.py`` checks them against each other and against the golden vectors.
"""
from __future__ import annotations

import itertools
import math
from typing import Iterator, List, Sequence, Tuple

import numpy as np

Cls = Tuple[int, ...]


# --------------------------------------------------------------------------------------------------
# permutation classes
# --------------------------------------------------------------------------------------------------
def _partitions_desc(n: int, max_part: int) -> Iterator[Cls]:
    if n == 0:
        yield ()
        return
    for first in range(min(n, max_part), 0, -1):
        for rest in _partitions_desc(n - first, first):
            yield (first,) + rest


def perm_classes(rank: int) -> List[Cls]:
    """Partitions of ``rank`` in descending lexicographic order: (r,), (r-1,1), ..., (1,)*r.

    This is the key order of ``PermClsSymmetricTensor._data`` and of ``perm_classes``
    (``symtensor/utils.py:839-856, 1000-1002``; ``symtensor/permcls_symtensor.py:596, 666-667``).
    Rank 0 has the single class ``()``.
    """
    return list(_partitions_desc(rank, rank))


def class_runs(cls: Cls) -> List[Tuple[int, int]]:
    """Maximal runs of equal multiplicity: [(multiplicity, run length g), ...] in class order."""
    return [(m, len(list(g))) for m, g in itertools.groupby(cls)]


def permclass_size(cls: Cls, dim: int) -> int:
    """Number of stored components of a class: d!/(d-l)!/prod(g_j!)  (``symtensor/utils.py:925-933``).

    Exact integer arithmetic (the reference divides floats and asserts integrality).
    """
    l = len(cls)
    if l > dim:
        return 0
    return math.perm(dim, l) // math.prod(math.factorial(g) for _, g in class_runs(cls))


def permclass_multiplicity(cls: Cls) -> int:
    """gamma = r!/prod(m_k!): how often one stored component occurs in the dense tensor
    (``symtensor/utils.py:207-223, 760-776``)."""
    return math.factorial(sum(cls)) // math.prod(math.factorial(m) for m in cls)


def indep_size(rank: int, dim: int) -> int:
    """C(d+r-1, r) (``symtensor/base.py:833-844``), written as in ``tests/test_utils.py:82`` so that
    rank 0 gives 1 even for dim 0."""
    return math.prod(range(dim, dim + rank)) // math.factorial(rank)


def get_permclass(index: Sequence[int]) -> Cls:
    """Class of a dense multi-index: value counts, sorted descending (``symtensor/utils.py:878-889``)."""
    return tuple(sorted((len(list(g)) for _, g in itertools.groupby(sorted(index))), reverse=True))


def index_representative(index: Sequence[int]) -> Tuple[int, ...]:
    """Representative of an index class: group equal values, order groups by count (descending),
    ties keep ascending value (stable sort)  (``symtensor/permcls_symtensor.py:375-381``)."""
    groups = [(v, len(list(g))) for v, g in itertools.groupby(sorted(index))]
    groups.sort(key=lambda t: -t[1])  # stable: ascending value inside equal counts
    out: Tuple[int, ...] = ()
    for v, c in groups:
        out += (v,) * c
    return out


def class_label(cls: Cls) -> str:
    """(2,1,1) -> 'iijk'  (``symtensor/utils.py:699, 735-744``)."""
    letters = "ijklmnabcdefghopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    return "".join(s * c for s, c in zip(letters, cls))


# --------------------------------------------------------------------------------------------------
# storage order inside a class -- generator semantics of the reference
# --------------------------------------------------------------------------------------------------
def sigma_index_values(cls: Cls, dim: int) -> Iterator[Tuple[int, ...]]:
    """Distinct values (v_0..v_{l-1}) of every stored component of ``cls`` in storage order.

    Restates ``_sub_σindex_iter`` / ``σindex_iter`` (``symtensor/permcls_symtensor.py:288-347``):
    positions are filled left to right, each loop ascending; position k may not reuse a value; if
    multiplicity m_k == m_{k-1} then additionally v_k > v_{k-1}.
    """
    l = len(cls)
    if l == 0:
        yield ()
        return
    if l > dim:
        return
    vals = [0] * l

    def fill(k: int, used: frozenset) -> Iterator[Tuple[int, ...]]:
        lo = vals[k - 1] + 1 if (k > 0 and cls[k] == cls[k - 1]) else 0
        for v in range(lo, dim):
            if v in used:
                continue
            vals[k] = v
            if k == l - 1:
                yield tuple(vals)
            else:
                yield from fill(k + 1, used | {v})

    yield from fill(0, frozenset())


def sigma_index_iter(cls: Cls, dim: int) -> Iterator[Tuple[int, ...]]:
    """Representative multi-index (length rank) of every stored component, in storage order."""
    for vals in sigma_index_values(cls, dim):
        idx: Tuple[int, ...] = ()
        for v, m in zip(vals, cls):
            idx += (v,) * m
        yield idx


# --------------------------------------------------------------------------------------------------
# storage order inside a class -- closed form (what the CUDA enumerator implements)
# --------------------------------------------------------------------------------------------------
def comb_lexrank(c: Sequence[int], n: int) -> int:
    """Lexicographic rank of the increasing combination ``c`` among the C(n, len(c)) combinations of
    range(n) (same formula as ``symtensor/flat_symtensor.py:28-36``)."""
    g = len(c)
    r = math.comb(n, g) - 1
    for k, ck in enumerate(reversed(c)):
        r -= math.comb(n - 1 - ck, k + 1)
    return r


def comb_lexunrank(rank: int, n: int, g: int) -> Tuple[int, ...]:
    """Inverse of ``comb_lexrank``."""
    out = []
    v = 0
    for i in range(g):
        while True:
            cnt = math.comb(n - 1 - v, g - 1 - i)  # combos that put v at position i
            if rank < cnt:
                break
            rank -= cnt
            v += 1
        out.append(v)
        v += 1
    return tuple(out)


def permcls_rank(cls: Cls, dim: int, vals: Sequence[int]) -> int:
    """Position of the component with distinct values ``vals`` (class order) inside its class.

    Mixed radix over the runs of equal multiplicity; run j is an increasing g_j-combination of the
    R_j values not used by earlier runs, relabelled by their order among the unused values
    (equivalent to enumerating ``σindex_iter``).
    """
    pos = 0
    used: List[int] = []
    k = 0
    for _, g in class_runs(cls):
        run = vals[k:k + g]
        n = dim - len(used)
        rel = [v - sum(1 for u in used if u < v) for v in run]
        pos = pos * math.comb(n, g) + comb_lexrank(rel, n)
        used.extend(run)
        k += g
    return pos


def permcls_unrank(cls: Cls, dim: int, pos: int) -> Tuple[int, ...]:
    """Inverse of ``permcls_rank``: distinct values (class order) of the component at ``pos``."""
    runs = class_runs(cls)
    radices = []
    n = dim
    for _, g in runs:
        radices.append(math.comb(n, g))
        n -= g
    digits = []
    for rdx in reversed(radices):
        digits.append(pos % rdx)
        pos //= rdx
    digits.reverse()
    vals: List[int] = []
    for (_, g), dig in zip(runs, digits):
        n = dim - len(vals)
        rel = comb_lexunrank(dig, n, g)
        used = sorted(vals)
        for v in rel:
            for u in used:  # undo the relabelling: skip values taken by earlier runs
                if v >= u:
                    v += 1
            vals.append(v)
    return tuple(vals)


def rank_of_index(index: Sequence[int], dim: int) -> Tuple[Cls, int]:
    """Dense multi-index -> (class, position): what ``_convert_dense_index`` looks up in the
    position registry (``symtensor/permcls_symtensor.py:422-479``)."""
    rep = index_representative(index)
    cls = get_permclass(index)
    vals = []
    k = 0
    for m in cls:
        vals.append(rep[k])
        k += m
    return cls, permcls_rank(cls, dim, vals)


# --------------------------------------------------------------------------------------------------
# vectorised enumeration (NumPy) -- used by the packed oracle
# --------------------------------------------------------------------------------------------------
def _combinations_array(n: int, g: int) -> np.ndarray:
    """All increasing g-combinations of range(n), lexicographic, as an int32 array [C(n,g), g]."""
    cnt = math.comb(n, g)
    if g == 0:
        return np.zeros((1, 0), dtype=np.int32)
    if cnt == 0:
        return np.zeros((0, g), dtype=np.int32)
    if g == 1:
        return np.arange(n, dtype=np.int32)[:, None]
    # first element f = 0..n-g; tail = combinations of the larger values
    parts = []
    for f in range(n - g + 1):
        tail = _combinations_array(n - f - 1, g - 1) + (f + 1)
        head = np.full((tail.shape[0], 1), f, dtype=np.int32)
        parts.append(np.concatenate([head, tail], axis=1))
    return np.concatenate(parts, axis=0)


def class_values(cls: Cls, dim: int) -> np.ndarray:
    """int32 array [size, l]: the distinct values of every stored component, in storage order."""
    l = len(cls)
    if l > dim:
        return np.zeros((0, l), dtype=np.int32)
    out = np.zeros((1, 0), dtype=np.int32)
    for _, g in class_runs(cls):
        n = dim - out.shape[1]
        rel = _combinations_array(n, g)  # [c, g]
        nh, nc = out.shape[0], rel.shape[0]
        head = np.repeat(out, nc, axis=0)  # head-major, combination fastest
        vals = np.tile(rel, (nh, 1))
        if head.shape[1]:
            srt = np.sort(head, axis=1)
            for j in range(srt.shape[1]):
                vals = vals + (vals >= srt[:, j:j + 1])
        out = np.concatenate([head, vals.astype(np.int32)], axis=1)
    return out


def class_repindex(cls: Cls, dim: int) -> np.ndarray:
    """int32 array [size, rank]: representative multi-indices in storage order."""
    vals = class_values(cls, dim)
    return np.repeat(vals, np.asarray(cls, dtype=np.int64), axis=1) if len(cls) else vals


# --------------------------------------------------------------------------------------------------
# flat format
# --------------------------------------------------------------------------------------------------
def flat_rank(dim: int, idx: Sequence[int]) -> int:
    """Position of the sorted multi-index in ``combinations_with_replacement(range(dim), r)`` order
    (``symtensor/flat_symtensor.py:39-50``)."""
    r = len(idx)
    pos = math.comb(dim + r - 1, r) - 1
    for k, ck in enumerate(reversed(idx)):
        pos -= math.comb(dim - 1 + k - ck, k + 1)
    return pos


def flat_unrank(dim: int, rank: int, pos: int) -> Tuple[int, ...]:
    """Inverse of ``flat_rank`` (multiset i_1<=...<=i_r  <->  strict combination i_k + k)."""
    c = comb_lexunrank(pos, dim + rank - 1, rank)
    return tuple(ck - k for k, ck in enumerate(c))


def flat_indices(rank: int, dim: int) -> np.ndarray:
    """int32 [N, rank] sorted multi-indices in flat storage order
    (``symtensor/flat_symtensor.py:219-220``)."""
    c = _combinations_array(dim + rank - 1, rank)
    return (c - np.arange(rank, dtype=np.int32)[None, :]).astype(np.int32)


def flat_multiplicity(idx: Sequence[int]) -> int:
    """r!/prod(n_v!) for a sorted multi-index (``symtensor/flat_symtensor.py:59-74``)."""
    return math.factorial(len(idx)) // math.prod(
        math.factorial(len(list(g))) for _, g in itertools.groupby(idx))
