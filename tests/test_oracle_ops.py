"""The floating-point oracles against outputs of the unmodified reference (tests/golden/op_goldens.*).

Tolerance: 1e-12 relative (fp64), the north-star bar; the reference's own results are only defined up to
summation order, and its rank-8 symmetrization carries ~2.5e-12, hence 1e-11 for n >= 7.
"""
import numpy as np
import pytest

from oracle import dense_oracle as do
from oracle import index_oracle as io
from oracle import packed_oracle as po

RTOL = 1e-12


def assert_packed_close(got, ref, rank, dim, rtol=RTOL):
    scale = max(float(np.max(np.abs(v))) for v in ref.values() if np.size(v)) or 1.0
    for cls in io.perm_classes(rank):
        if len(cls) > dim:
            continue
        g, r = np.asarray(got[cls]), np.asarray(ref[cls])
        assert g.shape == r.shape, (cls, g.shape, r.shape)
        assert np.max(np.abs(g - r)) <= rtol * scale, (cls, np.max(np.abs(g - r)) / scale)


def test_vec_against_reference(goldens):
    for c in goldens.cases("vec"):
        A, x, ref = goldens.packed(c["tag"] + ".A"), goldens.ops[c["tag"] + ".x"], float(goldens.ops[c["tag"] + ".out"])
        got_p = po.contract_all_indices_with_vector(A, c["rank"], c["dim"], x)
        terms = po.contract_all_indices_with_vector({k: np.abs(v) for k, v in A.items()}, c["rank"], c["dim"], np.abs(x))
        assert abs(got_p - ref) <= RTOL * terms, c
        if c["dim"] ** c["rank"] <= 5000:
            got_d = do.contract_all_indices_with_vector(A, c["rank"], c["dim"], x)
            assert abs(got_d - ref) <= RTOL * terms, c


def test_config1_against_reference(goldens):
    """BASELINE config 1 (rank 4, dim 50, fp64), value produced by the reference itself in ~7 s."""
    c = goldens.config1
    rng = np.random.default_rng(c["seed"])
    data = {cls: rng.uniform(0.5, 1.5, io.permclass_size(cls, 50)) for cls in io.perm_classes(4)}
    x = rng.uniform(0.5, 1.5, 50) / np.sqrt(50)
    got = po.contract_all_indices_with_vector(data, 4, 50, x)
    assert abs(got - c["result"]) <= RTOL * abs(c["result"])
    assert sum(v.size for v in data.values()) == c["packed_components"] == 292825


def test_flat_vec_against_reference(goldens):
    for c in goldens.cases("flatvec"):
        A, x, ref = goldens.ops[c["tag"] + ".A"], goldens.ops[c["tag"] + ".x"], float(goldens.ops[c["tag"] + ".out"])
        assert abs(po.contract_vec_flat(A, c["rank"], c["dim"], x) - ref) <= RTOL * abs(ref)
        # flat <-> permcls re-ordering is value preserving
        pc = po.flat_to_permcls(A, c["rank"], c["dim"])
        assert abs(po.contract_all_indices_with_vector(pc, c["rank"], c["dim"], x) - ref) <= RTOL * abs(ref)
        assert (po.permcls_to_flat(pc, c["rank"], c["dim"]) == A).all()


def test_mat_against_reference(goldens):
    for c in goldens.cases("mat"):
        A, W, ref = goldens.packed(c["tag"] + ".A"), goldens.ops[c["tag"] + ".W"], goldens.packed(c["tag"] + ".out")
        assert_packed_close(po.contract_all_indices_with_matrix(A, c["rank"], c["dim"], W), ref, c["rank"], c["dim"], 1e-11)
        assert_packed_close(do.contract_all_indices_with_matrix(A, c["rank"], c["dim"], W), ref, c["rank"], c["dim"], 1e-11)


def test_tensordot_against_reference(goldens):
    for c in goldens.cases("tensordot"):
        A, B, ref = goldens.packed(c["tag"] + ".A"), goldens.packed(c["tag"] + ".B"), goldens.packed(c["tag"] + ".out")
        got, n = po.tensordot(A, c["ra"], B, c["rb"], c["dim"], c["k"])
        assert n == c["out_rank"]
        if n == 0:
            assert c["out_dim"] == 1
            assert abs(float(got[()]) - float(ref[()])) <= 1e-11 * abs(float(ref[()]))
            continue
        assert_packed_close(got, ref, n, c["dim"], 1e-11)
        got_d, n_d = do.tensordot(A, c["ra"], B, c["rb"], c["dim"], c["k"])
        assert_packed_close(got_d, ref, n, c["dim"], 1e-11)


def test_outer_against_reference(goldens):
    for c in goldens.cases("outer"):
        A, B, ref = goldens.packed(c["tag"] + ".A"), goldens.packed(c["tag"] + ".B"), goldens.packed(c["tag"] + ".out")
        n = c["ra"] + c["rb"]
        assert_packed_close(po.outer(A, c["ra"], B, c["rb"], c["dim"]), ref, n, c["dim"], 1e-11)
        if c["dim"] ** n <= 20000 and n <= 6:
            assert_packed_close(do.outer(A, c["ra"], B, c["rb"], c["dim"]), ref, n, c["dim"], 1e-11)
    # e0 (x) e1 : off-diagonal 0.5, diagonal 0 (symtensor/testing/api.py:497-512)
    ref = goldens.packed("outer_e0e1.out")
    assert np.allclose(ref[(1, 1)], [0.5]) and np.allclose(ref[(2,)], [0.0, 0.0])
    got = po.outer({(1,): np.array([1.0, 0.0])}, 1, {(1,): np.array([0.0, 1.0])}, 1, 2)
    assert np.allclose(got[(1, 1)], [0.5]) and np.allclose(got[(2,)], [0.0, 0.0])


def test_algebraic_identities_beyond_reference_reach():
    """Identities used for full-size verification, here at sizes the dense path cannot do."""
    rng = np.random.default_rng(7)
    d, ra, rb = 9, 3, 3
    A = {c: rng.uniform(0.5, 1.5, io.permclass_size(c, d)) for c in io.perm_classes(ra)}
    B = {c: rng.uniform(0.5, 1.5, io.permclass_size(c, d)) for c in io.perm_classes(rb)}
    x = rng.uniform(0.5, 1.5, d)
    C = po.outer(A, ra, B, rb, d)
    lhs = po.contract_all_indices_with_vector(C, ra + rb, d, x)
    rhs = po.contract_all_indices_with_vector(A, ra, d, x) * po.contract_all_indices_with_vector(B, rb, d, x)
    assert abs(lhs - rhs) <= 1e-12 * abs(rhs)
    # tensordot(A, B, 0) == outer(A, B)   (symtensor/testing/api.py:521-522)
    T0, n0 = po.tensordot(A, ra, B, rb, d, 0)
    assert n0 == 6
    for c in C:
        assert np.allclose(T0[c], C[c], rtol=1e-13, atol=0)
    # vec(mat(A, W), y) == vec(A, W y)
    W, y = rng.uniform(0.5, 1.5, (d, d)) / d, rng.uniform(0.5, 1.5, d)
    M = po.contract_all_indices_with_matrix(A, ra, d, W)
    lhs = po.contract_all_indices_with_vector(M, ra, d, y)
    rhs = po.contract_all_indices_with_vector(A, ra, d, W @ y)
    assert abs(lhs - rhs) <= 1e-12 * abs(rhs)
