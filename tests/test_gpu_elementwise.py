"""SURVEY.md 8f.2-8f.4 on the GPU: elementwise ufuncs and comparisons on the packed buffer (symtensor/base.py:1146-1362,
1521-1684), partial indexing A[i] (symtensor/permcls_symtensor.py:750-781) and contract_tensor_list (symtensor/symalg.py:556-642),
against NumPy on the same packed data / the dense oracle / outputs of the unmodified reference (tests/golden/next_goldens)."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle as do
from oracle import index_oracle as io

import symtensor_b200 as st

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rand_packed(rank, dim, rng, positive=False):
    return {c: (rng.uniform(0.5, 1.5, io.permclass_size(c, dim)) if positive else rng.standard_normal(io.permclass_size(c, dim)))
            for c in io.perm_classes(rank)}


def classes_equal(T, ref, rtol=0.0):
    got = T.to_numpy_dict()
    for c, r in ref.items():
        if io.permclass_size(c, T.dim) == 0 and T.rank:
            continue
        g = np.asarray(got[c], dtype=np.float64).reshape(np.shape(r))
        assert np.allclose(g, r, rtol=rtol, atol=0, equal_nan=True), c


@pytest.mark.parametrize("cls_name", ["PermClsTorchSymmetricTensor", "FlatSymmetricTensor"])
@pytest.mark.parametrize("rank,dim", [(3, 7), (4, 5), (2, 40), (1, 9)])
def test_elementwise_ufuncs_match_numpy_on_the_packed_data(cls_name, rank, dim):
    rng = np.random.default_rng(rank * 10 + dim)
    A, B = rand_packed(rank, dim, rng, True), rand_packed(rank, dim, rng, True)
    if cls_name == "FlatSymmetricTensor":
        from oracle import packed_oracle as po
        fa, fb = po.permcls_to_flat(A, rank, dim), po.permcls_to_flat(B, rank, dim)
        TA, TB = st.FlatSymmetricTensor(rank, dim, fa, device=DEV), st.FlatSymmetricTensor(rank, dim, fb, device=DEV)
        check = lambda T, f, rtol=0.0: np.allclose(T.packed.cpu().numpy(), f(fa, fb), rtol=rtol, atol=0)  # noqa: E731
    else:
        TA = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=A, device=DEV)
        TB = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=B, device=DEV)

        def check(T, f, rtol=0.0):
            classes_equal(T, {c: f(A[c], B[c]) for c in A}, rtol)
            tab = T.class_table  # the alignment padding stays zero whatever the ufunc does to a zero
            for s, o, nxt in zip(tab.sizes, tab.offsets, list(tab.offsets[1:])):
                assert float(T.packed[o + s:nxt].abs().sum()) == 0.0
            return True
    assert check(TA + TB, lambda a, b: a + b) and check(np.subtract(TA, TB), lambda a, b: a - b)
    assert check(TA * TB, lambda a, b: a * b) and check(TA / TB, lambda a, b: a / b)
    assert check(2.5 * TA, lambda a, b: 2.5 * a) and check(TA - 1.0, lambda a, b: a - 1.0) and check(1.0 / TA, lambda a, b: 1.0 / a)
    ulp = 4.5e-16  # exp / log / pow come from the CUDA math library: correctly rounded to within an ulp or two
    assert check(np.exp(TA), lambda a, b: np.exp(a), ulp) and check(np.sqrt(TA), lambda a, b: np.sqrt(a)) and check(-TA, lambda a, b: -a)
    assert check(np.maximum(TA, TB), np.maximum) and check(TA ** 2, lambda a, b: a ** 2) and check(np.log(TA), lambda a, b: np.log(a), 1e-14)
    assert check(TA ** 2.5, lambda a, b: a ** 2.5, ulp) and check(np.power(TA, TB), np.power, ulp) and check(abs(-TA), lambda a, b: a)
    C = TA.copy()
    C += TB
    C *= 0.5
    assert check(C, lambda a, b: (a + b) * 0.5)
    assert type(TA + TB) is type(TA) and (TA + TB).rank == rank
    # comparisons (NEP 18): array_equal / allclose -> bool, isclose -> a 1.0 / 0.0 tensor
    assert np.array_equal(TA, TA.copy()) and not np.array_equal(TA, TB)
    assert np.allclose(TA, TA * (1 + 1e-9)) and not np.allclose(TA, TA * 1.01)
    assert np.allclose(TA, TA * 1.01, rtol=0.05)
    M = np.isclose(TA, TB)
    assert type(M) is type(TA) and float(M.packed.sum()) == 0.0
    assert float(np.isclose(TA, TA).packed.sum()) == TA.indep_size
    other = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim + 1, data=1.0, device=DEV) if cls_name != "FlatSymmetricTensor" else \
        st.FlatSymmetricTensor(rank, dim + 1, 1.0, device=DEV)
    assert not np.array_equal(TA, other)
    with pytest.raises(TypeError):
        TA + other  # different shapes: NotImplemented from both sides


@pytest.mark.parametrize("rank,dim", [(3, 6), (4, 5), (5, 4), (2, 9)])
def test_partial_indexing_matches_the_dense_slice(rank, dim):
    """A[i] / A[i, j] / A[i, :, :] (symtensor/permcls_symtensor.py:750-781): the lower-rank tensor equals the dense slice."""
    rng = np.random.default_rng(rank * 100 + dim)
    A = rand_packed(rank, dim, rng)
    dense = do.todense(A, rank, dim)
    TA = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=A, device=DEV)
    from oracle import packed_oracle as po
    TF = st.FlatSymmetricTensor(rank, dim, po.permcls_to_flat(A, rank, dim), device=DEV)
    for key in [(dim - 1,), (0,), (1, 2)[:min(2, rank - 1)], tuple(range(rank - 1))]:
        for T in (TA, TF):
            B = T[key] if len(key) > 1 else T[key[0]]
            assert type(B) is type(T) and B.rank == rank - len(key) and B.dim == dim
            assert np.array_equal(B.todense().cpu().numpy(), dense[key])
    assert np.array_equal(TA[(1,) + (slice(None),) * (rank - 1)].todense().cpu().numpy(), dense[1])
    with pytest.raises(NotImplementedError):
        TA[(slice(0, 2),) + (0,) * (rank - 1)]


def test_contract_tensor_list_against_dense_einsum():
    """symtensor/symalg.py:556-642 and its test symtensor/testing/api.py:616-654 (rule='all'), plus the default rule the
    reference cannot run (NameError): the sum over the second half of the index range."""
    rng = np.random.default_rng(9)
    for dim in (2, 3, 4, 5):
        A = rand_packed(3, dim, rng)
        TA = st.PermClsTorchSymmetricTensor(rank=3, dim=dim, data=A, device=DEV)
        chis, chi_dense = [], np.zeros((dim,) * 3)
        for i in range(dim):
            X = rand_packed(2, dim, rng)
            chis.append(st.PermClsTorchSymmetricTensor(rank=2, dim=dim, data=X, device=DEV))
            chi_dense[i] = do.todense(X, 2, dim)
        Ad = do.todense(A, 3, dim)
        c1 = st.contract_tensor_list(TA, chis, n_times=1, rule="all")
        c2 = st.contract_tensor_list(TA, chis, n_times=2, rule="all")
        assert c1.rank == 4 and c2.rank == 5
        assert np.allclose(c1.todense().cpu().numpy(), do.symmetrize(np.einsum("ija,akl->ijkl", Ad, chi_dense)), rtol=1e-11, atol=1e-12)
        assert np.allclose(c2.todense().cpu().numpy(), do.symmetrize(np.einsum("iab,ajk,blm->ijklm", Ad, chi_dense, chi_dense)), rtol=1e-11, atol=1e-12)
        h = -(-dim // 2)
        c3 = st.contract_tensor_list(TA, chis, n_times=1)  # rule='second_half'
        assert np.allclose(c3.todense().cpu().numpy(), do.symmetrize(np.einsum("ija,akl->ijkl", Ad[:, :, h:], chi_dense[h:])), rtol=1e-11, atol=1e-12)
    v = st.PermClsTorchSymmetricTensor(rank=1, dim=3, data={(1,): np.array([1.0, -2.0, 0.5])}, device=DEV)
    chis = [st.PermClsTorchSymmetricTensor(rank=2, dim=3, data=rand_packed(2, 3, rng), device=DEV) for _ in range(3)]
    c = st.contract_tensor_list(v, chis)
    want = sum(w * x.todense().cpu().numpy() for w, x in zip([1.0, -2.0, 0.5], chis))
    assert np.allclose(c.todense().cpu().numpy(), want)
    with pytest.raises(ValueError):
        st.contract_tensor_list(v, chis[:2])
