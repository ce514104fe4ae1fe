"""Host-side index bookkeeping, served by the C-ABI's exact-integer class tables.

Mirrors the names of the reference's ``symtensor.utils`` helpers for the hot path
(``_perm_classes`` utils.py:1000-1002, ``_get_permclass_size`` :925-933, ``get_permclass_multiplicity`` :760-776,
``_get_permclass`` :878-889, label<->counts :708-750) and ``permcls_symtensor.get_index_representative``
(:375-381) / ``_convert_dense_index`` (:448-479).
"""
from __future__ import annotations

import ctypes
import functools
from dataclasses import dataclass
from typing import List, Sequence, Tuple

from . import _cabi
from ._cabi import c_i64, check, lib

Cls = Tuple[int, ...]
INDEX_LETTERS = "ijklmnabcdefghopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"  # symtensor/utils.py:699


@dataclass(frozen=True)
class ClassTable:
    rank: int
    dim: int
    classes: Tuple[Cls, ...]     # `_perm_classes(rank)` order
    sizes: Tuple[int, ...]       # stored components per class
    mults: Tuple[int, ...]       # gamma per class
    offsets: Tuple[int, ...]     # start of each class in the packed buffer (+ padded total as last entry)

    @property
    def total(self) -> int:
        return self.offsets[-1]

    @property
    def ncls(self) -> int:
        return len(self.classes)

    def index(self, cls: Cls) -> int:
        return self.classes.index(tuple(cls))


@functools.lru_cache(maxsize=None)
def class_table(rank: int, dim: int) -> ClassTable:
    n = lib.st_num_classes(rank)
    if n < 0:
        check(_cabi.ST_ERR_INVALID)
    parts = (ctypes.c_int32 * (n * max(rank, 1)))()
    nparts = (ctypes.c_int32 * n)()
    sizes = (ctypes.c_int64 * n)()
    mults = (ctypes.c_int64 * n)()
    offs = (ctypes.c_int64 * (n + 1))()
    check(lib.st_class_table(rank, c_i64(dim), parts, nparts, sizes, mults, offs))
    classes = tuple(tuple(parts[c * rank:c * rank + nparts[c]]) for c in range(n))
    return ClassTable(rank, dim, classes, tuple(sizes), tuple(mults), tuple(offs))


def perm_classes(rank: int) -> List[Cls]:
    return list(class_table(rank, 1).classes)


def permclass_size(cls: Cls, dim: int) -> int:
    t = class_table(sum(cls), dim)
    return t.sizes[t.index(cls)]


def permclass_multiplicity(cls: Cls) -> int:
    t = class_table(sum(cls), 1)
    return t.mults[t.index(cls)]


def indep_size(rank: int, dim: int) -> int:
    out = c_i64()
    check(lib.st_indep_size(rank, c_i64(dim), ctypes.byref(out)))
    return out.value


def permclass_counts_to_label(counts: Sequence[int]) -> str:
    return "".join(s * c for s, c in zip(INDEX_LETTERS, counts))


def permclass_label_to_counts(label: str) -> Cls:
    return tuple(sorted((label.count(s) for s in set(label)), reverse=True))


def get_permclass(index: Sequence) -> Cls:
    counts = {}
    for v in index:
        counts[v] = counts.get(v, 0) + 1
    return tuple(sorted(counts.values(), reverse=True))


def convert_dense_index(rank: int, dim: int, index: Sequence[int]) -> Tuple[Cls, int]:
    """Dense multi-index -> (class, position inside the class)."""
    if len(index) != rank:
        raise IndexError("Partial indexing (fewer indices than the rank) is not supported on packed storage.")
    arr = (ctypes.c_int32 * max(rank, 1))(*[int(i) for i in index])
    cls, pos = ctypes.c_int32(), c_i64()
    try:
        check(lib.st_host_permcls_rank(rank, c_i64(dim), arr, ctypes.byref(cls), ctypes.byref(pos)))
    except ValueError as e:
        raise IndexError(str(e)) from None
    return class_table(rank, dim).classes[cls.value], pos.value


def index_of(rank: int, dim: int, cls: Cls, pos: int) -> Tuple[int, ...]:
    """(class, position) -> representative multi-index (inverse of ``convert_dense_index``)."""
    out = (ctypes.c_int32 * max(rank, 1))()
    check(lib.st_host_permcls_unrank(rank, c_i64(dim), class_table(rank, dim).index(cls), c_i64(pos), out))
    return tuple(out[:rank])


def get_index_representative(index: Sequence[int]) -> Tuple[int, ...]:
    rank = len(index)
    dim = max(index) + 1 if rank else 1
    cls, pos = convert_dense_index(rank, dim, index)
    return index_of(rank, dim, cls, pos)


def flat_rank(dim: int, index: Sequence[int]) -> int:
    rank = len(index)
    arr = (ctypes.c_int32 * max(rank, 1))(*[int(i) for i in index])
    pos = c_i64()
    check(lib.st_host_flat_rank(rank, c_i64(dim), arr, ctypes.byref(pos)))
    return pos.value


def flat_unrank(rank: int, dim: int, pos: int) -> Tuple[int, ...]:
    out = (ctypes.c_int32 * max(rank, 1))()
    check(lib.st_host_flat_unrank(rank, c_i64(dim), c_i64(pos), out))
    return tuple(out[:rank])
