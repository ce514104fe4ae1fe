"""The rows either side of the hot path on the GPU against the unmodified reference's outputs
(tests/golden/next_goldens.*, oracle/gen_golden_next.py) and the dense oracle: add.outer / subtract.outer
(symtensor/symalg.py:294-316) through the reference-facing registry, construction from dense / todense.

Tolerance: 1e-12 of the per-component mean of |terms| (fp64); 1e-5 (fp32, against the fp64 oracle on the rounded inputs)."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle as do
from oracle import index_oracle as io

import symtensor_b200 as st
from symtensor_b200._cabi import OUTER_ADD, c_i64, lib
from test_oracle_next import NextGoldens, outer_op_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nxt():
    return NextGoldens()


def absd(d):
    return {k: np.abs(v) for k, v in d.items()}


def check_classes(T, ref, scale, rtol):
    got = T.to_numpy_dict()
    for cls, r in ref.items():
        if len(cls) > T.dim:
            continue
        g = np.asarray(got[cls], dtype=np.float64)
        assert np.all(np.abs(g - r) <= rtol * np.maximum(scale[cls], 1e-300)), (cls, np.max(np.abs(g - r)))


@pytest.mark.parametrize("opname", ["add", "subtract"])
def test_add_subtract_outer_match_the_reference(nxt, opname):
    op = getattr(st, opname) if hasattr(st, opname) else getattr(st.symalg, opname)
    for c in nxt.cases(opname + "_outer"):
        a, b, ref = nxt.packed(c["tag"] + ".A"), nxt.packed(c["tag"] + ".B"), nxt.packed(c["tag"] + ".out")
        ra, rb, d = c["ra"], c["rb"], c["dim"]
        A = st.PermClsTorchSymmetricTensor(rank=ra, dim=d, data=dict(a), device=DEV)
        B = st.PermClsTorchSymmetricTensor(rank=rb, dim=d, data=dict(b), device=DEV)
        C = op.outer(A, B)
        assert isinstance(C, st.PermClsTorchSymmetricTensor) and C.rank == ra + rb and C.dim == d
        scale = outer_op_oracle("add", absd(a), ra, absd(b), rb, d)
        check_classes(C, ref, scale, 1e-12)
        # flat operands give a flat result with the same components
        FA = st.FlatSymmetricTensor(ra, d, data=A.todense(), device=DEV)
        FB = st.FlatSymmetricTensor(rb, d, data=B.todense(), device=DEV)
        FC = op.outer(FA, FB)
        assert isinstance(FC, st.FlatSymmetricTensor)
        assert torch.allclose(FC.todense(), C.todense(), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("ra,rb,dim", [(2, 2, 9), (3, 2, 7), (4, 1, 6), (4, 4, 4), (5, 3, 3), (2, 4, 5)])
def test_add_subtract_outer_against_the_dense_oracle(ra, rb, dim):
    rng = np.random.default_rng(100 * ra + 10 * rb + dim)
    a = {c: rng.standard_normal(io.permclass_size(c, dim)) for c in io.perm_classes(ra)}
    b = {c: rng.standard_normal(io.permclass_size(c, dim)) for c in io.perm_classes(rb)}
    A = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=a, device=DEV)
    B = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=b, device=DEV)
    scale = outer_op_oracle("add", absd(a), ra, absd(b), rb, dim)
    for opname in ("add", "subtract"):
        C = getattr(st.symalg, opname).outer(A, B)
        check_classes(C, outer_op_oracle(opname, a, ra, b, rb, dim), scale, 1e-12)
    # antisymmetry of subtract, and add - subtract = 2 * mean of B's splits: (A + B) - (A - B) == 2 * (0 + B)
    S1, S2 = st.symalg.subtract.outer(A, B), st.symalg.subtract.outer(B, A)
    assert torch.allclose(S1.packed, -S2.packed, rtol=1e-13, atol=1e-13)
    # fp32 against the fp64 oracle on the rounded inputs
    a32, b32 = {k: v.astype(np.float32) for k, v in a.items()}, {k: v.astype(np.float32) for k, v in b.items()}
    A32 = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=a32, device=DEV)
    B32 = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=b32, device=DEV)
    C32 = st.symalg.add.outer(A32, B32)
    assert C32.torch_dtype == torch.float32
    ref32 = outer_op_oracle("add", {k: v.astype(np.float64) for k, v in a32.items()}, ra, {k: v.astype(np.float64) for k, v in b32.items()}, rb, dim)
    check_classes(C32, ref32, scale, 1e-5)


def test_outer_with_a_scalar_and_with_out_and_errors():
    rng = np.random.default_rng(3)
    a = {c: rng.standard_normal(io.permclass_size(c, 5)) for c in io.perm_classes(3)}
    A = st.PermClsTorchSymmetricTensor(rank=3, dim=5, data=a, device=DEV)
    s = st.PermClsTorchSymmetricTensor(rank=0, dim=1, data=np.float64(2.5), device=DEV)
    for opname, f in [("add", lambda t: t + 2.5), ("subtract", lambda t: t - 2.5), ("multiply", lambda t: t * 2.5)]:
        C = getattr(st.symalg, opname).outer(A, s)
        for cls, v in a.items():
            assert np.allclose(C.to_numpy_dict()[cls], f(v), rtol=1e-15)
        assert float(C.packed.sum()) == pytest.approx(sum(float(f(v).sum()) for v in a.values()), rel=1e-12)  # padding stays zero
    C = st.symalg.subtract.outer(s, A)
    for cls, v in a.items():
        assert np.allclose(C.to_numpy_dict()[cls], 2.5 - v, rtol=1e-15)
    out = st.PermClsTorchSymmetricTensor(rank=6, dim=5, device=DEV)
    res = st.symalg.add.outer(A, A, out=out)
    assert res is out
    check_classes(out, outer_op_oracle("add", a, 3, a, 3, 5), outer_op_oracle("add", absd(a), 3, absd(a), 3, 5), 1e-12)
    B = st.PermClsTorchSymmetricTensor(rank=2, dim=4, device=DEV)
    with pytest.raises(TypeError):  # different dims: every implementation returns NotImplemented (symalg.py:305-308)
        st.symalg.add.outer(A, B)
    buf = torch.zeros(64, dtype=torch.float64, device=DEV)
    assert lib.st_outer_op_f64(9, 1, 1, c_i64(4), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), c_i64(0), c_i64(32), None) != 0  # unknown op
    assert lib.st_outer_op_f64(OUTER_ADD, 1, 1, c_i64(4), None, buf.data_ptr(), buf.data_ptr(), c_i64(0), c_i64(32), None) != 0  # null operand


def test_dense_construction_and_todense_match_the_reference(nxt):
    for c in nxt.cases("dense"):
        rank, dim, tag = c["rank"], c["dim"], c["tag"]
        a = nxt.packed(tag + ".A")
        dense = nxt.npz[tag + ".dense"]
        A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=dict(a), device=DEV)
        assert np.array_equal(A.todense().cpu().numpy(), dense)
        B = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=dense, device=DEV)
        for cls, r in nxt.packed(tag + ".repacked").items():
            assert np.array_equal(np.asarray(B.to_numpy_dict()[cls]), r)
        raw = nxt.npz[tag + ".raw"]
        S = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=raw, symmetrize=True, device=DEV)
        scale = do.repack(do.symmetrize(np.abs(raw)), rank, dim)
        check_classes(S, nxt.packed(tag + ".symmetrized"), scale, 1e-13)
        with pytest.raises(ValueError, match="not symmetric"):  # the reference rejected `raw` too (c["raw_rejected"])
            st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=raw, device=DEV)
        assert c["raw_rejected"]
