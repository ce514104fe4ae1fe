"""The integer oracle against the reference's golden vectors (CPU only).

Sources of truth: fixtures generated from the unmodified reference (tests/golden/index_goldens.*, made by
oracle/gen_golden.py) and the literal vectors in the reference's own tests, cited per test.
"""
import itertools
import math

import numpy as np
import pytest

from oracle import index_oracle as io
from conftest import key_cls


def test_perm_classes_order_sizes_multiplicities(goldens):
    meta = goldens.index_meta
    for r in range(0, 9):
        classes = io.perm_classes(r)
        assert [list(c) for c in classes] == meta["perm_classes"][str(r)]
        assert [io.permclass_multiplicity(c) for c in classes] == meta["multiplicities"][str(r)]
    for k, sizes in meta["sizes"].items():
        r, d = map(int, k.split(","))
        exact = [io.permclass_size(c, d) for c in io.perm_classes(r)]
        # The reference divides in floating point (symtensor/utils.py:930-933), so beyond 2**53 its value is
        # the correctly rounded double of the exact count; the oracle (and the CUDA tables) keep exact integers.
        assert [e if e < 2 ** 53 else int(float(e)) for e in exact] == sizes


def test_reference_literal_class_strings():
    # symtensor/testing/api.py:76-82
    assert [io.class_label(c) for c in io.perm_classes(2)] == ["ii", "ij"]
    assert [io.class_label(c) for c in io.perm_classes(4)] == ["iiii", "iiij", "iijj", "iijk", "ijkl"]
    assert [io.class_label(c) for c in io.perm_classes(5)] == ["iiiii", "iiiij", "iiijj", "iiijk", "iijjk", "iijkl", "ijklm"]


def test_reference_literal_sigma_index_iter():
    # symtensor/tests/test_permcls_numpy.py:159-163
    assert list(io.sigma_index_iter((2, 1), 3)) == [(0, 0, 1), (0, 0, 2), (1, 1, 0), (1, 1, 2), (2, 2, 0), (2, 2, 1)]
    assert list(io.sigma_index_iter((2, 2), 3)) == [(0, 0, 1, 1), (0, 0, 2, 2), (1, 1, 2, 2)]
    # symtensor/tests/test_permcls_numpy.py:172-176
    assert io.index_representative((2, 1, 2)) == (2, 2, 1)
    assert io.index_representative((5, 4, 3, 3, 2, 1)) == (3, 3, 1, 2, 4, 5)
    # storage positions, symtensor/testing/api.py:308-328: A[0,0,3] == A['iij'][2], A[1,2,3] == A['ijk'][6] at dim 4..
    assert io.rank_of_index((0, 0, 3), 4) == ((2, 1), 2)


def test_partition_identities():
    # symtensor/tests/test_utils.py:79-88
    for r in range(0, 9):
        for d in (0, 2, 10, 400):
            classes = io.perm_classes(r)
            assert sum(io.permclass_size(c, d) for c in classes) == io.indep_size(r, d)
            assert sum(io.permclass_size(c, d) * io.permclass_multiplicity(c) for c in classes) == d ** r


def test_storage_order_full(goldens):
    n = 0
    for k in goldens.index.files:
        if not k.startswith("sigma.r"):
            continue
        _, rs, ds, ck = k.split(".")
        r, d, cls = int(rs[1:]), int(ds[1:]), key_cls(ck)
        ref = goldens.index[k]
        rep = io.class_repindex(cls, d)
        assert rep.shape == ref.shape and (rep == ref).all(), k
        assert [tuple(t) for t in ref.tolist()] == list(io.sigma_index_iter(cls, d)), k
        vals = io.class_values(cls, d)
        for p in range(0, ref.shape[0], max(1, ref.shape[0] // 50)):
            v = tuple(vals[p].tolist())
            assert io.permcls_rank(cls, d, v) == p
            assert io.permcls_unrank(cls, d, p) == v
            assert io.rank_of_index(tuple(ref[p].tolist()), d) == (cls, p)
        n += 1
    assert n > 50


def test_storage_order_config1_sample(goldens):
    for cls in [(2, 1, 1), (1, 1, 1, 1), (3, 1), (2, 2)]:
        ck = "c" + "_".join(map(str, cls))
        pos = goldens.index[f"sigma_sample.r4.d50.{ck}.pos"]
        idx = goldens.index[f"sigma_sample.r4.d50.{ck}.idx"]
        rep = io.class_repindex(cls, 50)
        assert (rep[pos] == idx).all()
        w = (np.arange(rep.shape[0], dtype=np.int64)[:, None] + 1) * (rep.astype(np.int64) + 1)
        assert int(w.sum() % (2 ** 61 - 1)) == int(goldens.index[f"sigma_sample.r4.d50.{ck}.checksum"][0])
        for p, t in zip(pos.tolist(), idx.tolist()):
            assert io.rank_of_index(tuple(t), 50) == (cls, p)


def test_representatives(goldens):
    for idx, rep, cls in goldens.index_meta["representatives"]:
        assert list(io.index_representative(idx)) == rep
        assert list(io.get_permclass(idx)) == cls


def test_flat_order(goldens):
    for k in goldens.index.files:
        if k.startswith("flat.") and k.endswith(".idx"):
            _, rs, ds, _ = k.split(".")
            r, d = int(rs[1:]), int(ds[1:])
            idx = goldens.index[k]
            assert (io.flat_indices(r, d) == idx).all()
            ranks = goldens.index[k[:-3] + "rank"]
            mult = goldens.index[k[:-3] + "mult"]
            assert (ranks == np.arange(len(ranks))).all()
            for p, t in enumerate(idx.tolist()):
                assert io.flat_rank(d, t) == p
                assert io.flat_unrank(d, r, p) == tuple(t)
                assert io.flat_multiplicity(t) == mult[p]
        if k.startswith("flat_big.") and k.endswith(".idx"):
            _, rs, ds, _ = k.split(".")
            r, d = int(rs[1:]), int(ds[1:])
            ranks = goldens.index[k[:-3] + "rank"]
            for t, p in zip(goldens.index[k].tolist(), ranks.tolist()):
                assert io.flat_rank(d, t) == p
                assert io.flat_unrank(d, r, p) == tuple(t)


def test_closed_form_equals_generator_exhaustive():
    for r in range(1, 7):
        for d in range(0, 7):
            for cls in io.perm_classes(r):
                gen = list(io.sigma_index_values(cls, d))
                assert len(gen) == io.permclass_size(cls, d)
                arr = io.class_values(cls, d)
                assert [tuple(t) for t in arr.tolist()] == gen
                for p, v in enumerate(gen):
                    assert io.permcls_rank(cls, d, v) == p
                    assert io.permcls_unrank(cls, d, p) == v


def test_odometer_successor_reproduces_the_storage_order():
    """permcls_next_vals (the odometer step the multiply.outer kernel walks its runs of consecutive coordinates with) against
    the oracle's enumeration of sigma-index order: every class of ranks 1..8 at small dims, from several start positions."""
    import ctypes

    from symtensor_b200._cabi import c_i64, lib
    for rank, dim in [(1, 5), (2, 6), (3, 5), (4, 7), (5, 6), (6, 7), (8, 9), (4, 3), (6, 5)]:
        for ci, cls in enumerate(io.perm_classes(rank)):
            ref = io.class_repindex(cls, dim)
            n = len(ref)
            if n == 0:
                continue
            for start in sorted({0, n // 3, max(0, n - 5)}):
                want = ref[start:start + 400]
                buf = np.full((len(want) + 3, rank), -1, dtype=np.int32)
                got = lib.st_debug_permcls_successors(rank, c_i64(dim), ci, c_i64(start), c_i64(len(buf)), buf.ctypes.data)
                assert got == min(len(buf), n - start), (rank, dim, cls, start, got)
                assert np.array_equal(buf[:min(got, len(want))], want[:got]), (rank, dim, cls, start)


def test_row_walk_hands_every_coordinate_its_sorted_multi_index():
    """The warp-uniform row walk of the multiply.outer kernel (RowCursor, rowcursor_serve, row_component in st_common.cuh),
    replayed on the CPU by st_debug_rowwalk: every coordinate of the permcls layout gets the sorted multi-index the oracle's
    enumeration of the storage order gives it, padding included, for whole tensors and for spans that start mid-row."""
    from symtensor_b200._cabi import c_i64, lib
    from symtensor_b200 import combinatorics as comb
    for rank, dim in [(1, 5), (2, 6), (3, 5), (4, 7), (5, 6), (6, 7), (8, 9), (4, 3), (6, 5), (7, 4), (8, 3), (2, 40), (3, 37)]:
        tab = comb.class_table(rank, dim)
        want = np.full((tab.total, rank), -1, dtype=np.int32)
        for cls, off in zip(io.perm_classes(rank), tab.offsets):
            rep = io.class_repindex(cls, dim)
            if len(rep):
                want[off:off + len(rep)] = np.sort(rep, axis=1)
        for span in (tab.total, 1024, 96, 37):
            got = np.full((tab.total, rank), -7, dtype=np.int32)
            n = lib.st_debug_rowwalk(rank, c_i64(dim), c_i64(0), c_i64(tab.total), c_i64(span), got.ctypes.data)
            assert n == tab.total
            bad = np.nonzero((got != want).any(axis=1))[0]
            assert len(bad) == 0, (rank, dim, span, bad[:5], got[bad[:5]], want[bad[:5]])
        b, e = tab.total // 3, tab.total - 5
        got = np.full((e - b, rank), -7, dtype=np.int32)
        assert lib.st_debug_rowwalk(rank, c_i64(dim), c_i64(b), c_i64(e), c_i64(200), got.ctypes.data) == e - b
        assert np.array_equal(got, want[b:e]), (rank, dim, "range")
