"""GPU index-class enumerator (bulk rank / unrank kernels) against the reference's storage order: bit-exact.
Goldens come from the unmodified reference (tests/golden/index_goldens.npz); larger classes are checked against
the C oracle's generator-semantics enumeration and through size-independent properties at the full sizes of
BASELINE.json's configs (round trip, lexicographic monotonicity)."""
import numpy as np
import pytest
import torch

from conftest import key_cls
from oracle import c_oracle as co
from oracle import index_oracle as io

import symtensor_b200 as st
from symtensor_b200 import combinatorics as comb
from symtensor_b200._cabi import c_i64, check, lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def unrank(rank, dim, cls, begin, count):
    t = comb.class_table(rank, dim)
    out = torch.empty((count, rank), dtype=torch.int32, device=DEV)
    check(lib.st_permcls_unrank(rank, c_i64(dim), t.index(cls), c_i64(begin), c_i64(count), out.data_ptr(), None))
    torch.cuda.synchronize()
    return out


def rank_of(rank, dim, idx):
    n = idx.shape[0]
    cls = torch.empty(n, dtype=torch.int32, device=DEV)
    pos = torch.empty(n, dtype=torch.int64, device=DEV)
    check(lib.st_permcls_rank(rank, c_i64(dim), c_i64(n), idx.contiguous().data_ptr(), cls.data_ptr(), pos.data_ptr(), None))
    torch.cuda.synchronize()
    return cls, pos


def test_storage_order_matches_reference_goldens(goldens):
    n = 0
    for k in goldens.index.files:
        if not k.startswith("sigma.r"):
            continue
        _, rs, ds, ck = k.split(".")
        r, d, cls = int(rs[1:]), int(ds[1:]), key_cls(ck)
        ref = goldens.index[k].astype(np.int32)
        if ref.shape[0] == 0:
            continue
        got = unrank(r, d, cls, 0, ref.shape[0]).cpu().numpy()
        assert np.array_equal(got, ref), k
        # rank of arbitrarily permuted indices returns (class, position)
        perm = np.random.default_rng(n).permutation(r)
        c, p = rank_of(r, d, torch.as_tensor(ref[:, perm].copy(), device=DEV))
        assert (c.cpu().numpy() == comb.class_table(r, d).index(cls)).all()
        assert np.array_equal(p.cpu().numpy(), np.arange(ref.shape[0]))
        n += 1
    assert n > 50


def test_config1_classes_full_enumeration(goldens):
    """rank 4 dim 50 (config C1): every class, against the C oracle's literal σindex_iter loops and the
    checksums produced by the reference itself."""
    for cls in io.perm_classes(4):
        ref = co.class_repindex(cls, 50)
        got = unrank(4, 50, cls, 0, ref.shape[0]).cpu().numpy()
        assert np.array_equal(got, ref)
        ck = "c" + "_".join(map(str, cls))
        if f"sigma_sample.r4.d50.{ck}.checksum" in goldens.index.files:
            w = (np.arange(got.shape[0], dtype=np.int64)[:, None] + 1) * (got.astype(np.int64) + 1)
            assert int(w.sum() % (2 ** 61 - 1)) == int(goldens.index[f"sigma_sample.r4.d50.{ck}.checksum"][0])


@pytest.mark.parametrize("rank,dim", [(4, 200), (6, 64), (8, 40), (3, 1000)])
def test_full_size_round_trip_and_order(rank, dim):
    """At the configs' full sizes: unrank -> rank is the identity on sampled windows of every class, windows are
    lexicographically increasing in the distinct values, and class boundaries hold the expected first / last index."""
    t = comb.class_table(rank, dim)
    rng = np.random.default_rng(rank * 1000 + dim)
    for ci, (cls, size) in enumerate(zip(t.classes, t.sizes)):
        if size == 0:
            continue
        win = min(size, 4096)
        starts = sorted({0, size - win, *(int(s) for s in rng.integers(0, size - win + 1, 3))})
        for s0 in starts:
            idx = unrank(rank, dim, cls, s0, win)
            c, p = rank_of(rank, dim, idx)
            assert (c == ci).all()
            assert torch.equal(p, torch.arange(s0, s0 + win, device=DEV))
            # shuffled entries rank to the same place
            perm = torch.as_tensor(rng.permutation(rank), device=DEV)
            c2, p2 = rank_of(rank, dim, idx[:, perm].contiguous())
            assert torch.equal(p2, p) and (c2 == ci).all()
            # lexicographic order of the distinct values (take first entry of each multiplicity group)
            cols = np.cumsum((0,) + cls[:-1])
            v = idx[:, torch.as_tensor(cols, device=DEV)].to(torch.int64)
            key = torch.zeros(win, dtype=torch.float64, device=DEV)
            for j in range(v.shape[1]):
                key = key * dim + v[:, j].to(torch.float64)
            if dim ** len(cls) < 2 ** 52:
                assert (key[1:] > key[:-1]).all()
        first = unrank(rank, dim, cls, 0, 1).cpu().numpy()[0]
        assert tuple(first) == next(io.sigma_index_iter(cls, dim))


def test_large_class_matches_c_oracle_window():
    """(2,1,1) and (1,1,1,1) at dim 200 (config C2) against the C oracle on the whole class."""
    for cls in [(3, 1), (2, 2), (2, 1, 1)]:
        ref = co.class_repindex(cls, 200)
        got = unrank(4, 200, cls, 0, ref.shape[0]).cpu().numpy()
        assert np.array_equal(got, ref)
    ref = co.class_repindex((1, 1, 1, 1), 120)
    got = unrank(4, 120, (1, 1, 1, 1), 0, ref.shape[0]).cpu().numpy()
    assert np.array_equal(got, ref)


def test_flat_order(goldens):
    for k in goldens.index.files:
        if k.startswith("flat.") and k.endswith(".idx"):
            _, rs, ds, _ = k.split(".")
            r, d = int(rs[1:]), int(ds[1:])
            ref = goldens.index[k].astype(np.int32)
            out = torch.empty(ref.shape, dtype=torch.int32, device=DEV)
            check(lib.st_flat_unrank(r, c_i64(d), c_i64(0), c_i64(ref.shape[0]), out.data_ptr(), None))
            assert np.array_equal(out.cpu().numpy(), ref)
            pos = torch.empty(ref.shape[0], dtype=torch.int64, device=DEV)
            shuf = torch.as_tensor(ref[:, ::-1].copy(), device=DEV)
            check(lib.st_flat_rank(r, c_i64(d), c_i64(ref.shape[0]), shuf.data_ptr(), pos.data_ptr(), None))
            assert np.array_equal(pos.cpu().numpy(), np.arange(ref.shape[0]))
        if k.startswith("flat_big.") and k.endswith(".idx"):
            _, rs, ds, _ = k.split(".")
            r, d = int(rs[1:]), int(ds[1:])
            idx = torch.as_tensor(goldens.index[k].astype(np.int32), device=DEV)
            pos = torch.empty(idx.shape[0], dtype=torch.int64, device=DEV)
            check(lib.st_flat_rank(r, c_i64(d), c_i64(idx.shape[0]), idx.data_ptr(), pos.data_ptr(), None))
            assert np.array_equal(pos.cpu().numpy(), goldens.index[k[:-3] + "rank"])
    # round trip at the end of a large flat range
    n = comb.indep_size(4, 200)
    out = torch.empty((8192, 4), dtype=torch.int32, device=DEV)
    check(lib.st_flat_unrank(4, c_i64(200), c_i64(n - 8192), c_i64(8192), out.data_ptr(), None))
    pos = torch.empty(8192, dtype=torch.int64, device=DEV)
    check(lib.st_flat_rank(4, c_i64(200), c_i64(8192), out.data_ptr(), pos.data_ptr(), None))
    assert torch.equal(pos, torch.arange(n - 8192, n, device=DEV))
    assert out[-1].tolist() == [199, 199, 199, 199]


def test_invalid_arguments_return_errors():
    out = torch.empty((4, 3), dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError):
        check(lib.st_permcls_unrank(3, c_i64(4), 0, c_i64(2), c_i64(4), out.data_ptr(), None))  # class (3,) has 4 comps
    with pytest.raises(ValueError):
        check(lib.st_permcls_unrank(3, c_i64(4), 9, c_i64(0), c_i64(1), out.data_ptr(), None))
    bad = torch.tensor([[0, 1, 7]], dtype=torch.int32, device=DEV)
    c, p = rank_of(3, 4, bad)
    assert int(c[0]) == -1 and int(p[0]) == -1
