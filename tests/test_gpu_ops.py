"""multiply.outer, tensordot and contract_all_indices_with_matrix on the GPU, through the reference-facing API
(symtensor_b200.symalg, registered with ``@Cls.implements`` like a reference backend mixin) and the C-ABI, against:
the unmodified reference's outputs (tests/golden), the packed oracle on seeded inputs at sizes the reference cannot
reach, and the algebraic identities of SURVEY.md 8c.  Index maps (layout converters) are checked bit-exactly.

Tolerances (north-star): 1e-12 relative (fp64) / 1e-5 relative (fp32) of the sum of |terms| per component; fp32
parity is taken against the fp64 oracle on the up-cast inputs (the reference's own fp32 path raises)."""
import numpy as np
import pytest
import torch

from oracle import index_oracle as io
from oracle import packed_oracle as po

import symtensor_b200 as st
from symtensor_b200 import combinatorics as comb
from symtensor_b200 import ops
from symtensor_b200._cabi import c_i64, check, lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL64, RTOL32 = 1e-12, 1e-5


def rand_packed(rank, dim, rng, dist="pos", dtype=np.float64):
    out = {}
    for c in io.perm_classes(rank):
        n = io.permclass_size(c, dim)
        out[c] = (rng.uniform(0.5, 1.5, n) if dist == "pos" else rng.standard_normal(n)).astype(dtype)
    return out


def absd(d):
    return {k: np.abs(np.asarray(v, dtype=np.float64)) for k, v in d.items()}


def assert_classes_close(T, ref, scale, rtol):
    """T: GPU tensor; ref / scale: {class: array} (scale = same op on |inputs|, the per-component sum of |terms|)."""
    got = T.to_numpy_dict()
    for c, r in ref.items():
        if io.permclass_size(c, T.dim) == 0 and T.rank:
            continue
        g = np.asarray(got[c], dtype=np.float64).reshape(np.shape(r))
        s = np.maximum(np.asarray(scale[c], dtype=np.float64), 1e-300)
        assert np.all(np.abs(g - np.asarray(r, dtype=np.float64)) <= rtol * s), (c, np.max(np.abs(g - r) / s))


def full(data, rank, dim):
    return {k: np.broadcast_to(np.asarray(v, dtype=np.float64), (io.permclass_size(k, dim),)) if rank else np.asarray(v)
            for k, v in data.items()}


# ------------------------------------------------------------------------------------------------------------
def test_layout_converters_are_bit_exact():
    """permcls <-> flat re-ordering (the reference's σindex_iter / combinations_with_replacement orders)."""
    for rank, dim in [(1, 7), (2, 9), (3, 6), (4, 11), (5, 5), (6, 7), (8, 5)]:
        rng = np.random.default_rng(rank * 100 + dim)
        data = rand_packed(rank, dim, rng, "normal")
        A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device=DEV)
        flat = ops._flat_buffer(A, torch.float64)
        assert np.array_equal(flat.cpu().numpy(), po.permcls_to_flat(data, rank, dim))
        B = ops._wrap_flat_result(st.PermClsTorchSymmetricTensor, rank, dim, flat)
        assert torch.equal(B.packed, A.packed)  # values and zero padding


def test_outer_reference_goldens(goldens):
    for c in goldens.cases("outer"):
        A, B, ref = goldens.packed(c["tag"] + ".A"), goldens.packed(c["tag"] + ".B"), goldens.packed(c["tag"] + ".out")
        ra, rb, d = c["ra"], c["rb"], c["dim"]
        TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=d, data=dict(A), device=DEV)
        TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=d, data=dict(B), device=DEV)
        C = st.multiply.outer(TA, TB)
        assert isinstance(C, st.PermClsTorchSymmetricTensor) and C.rank == ra + rb and C.dim == d
        scale = po.outer(absd(full(A, ra, d)), ra, absd(full(B, rb, d)), rb, d)
        assert_classes_close(C, ref, scale, 1e-11)  # the reference's own (ra+rb)! symmetrization noise
        assert_classes_close(C, po.outer(full(A, ra, d), ra, full(B, rb, d), rb, d), scale, RTOL64)
    # e0 (x)_s e1: off-diagonal 0.5, diagonal 0 (symtensor/testing/api.py:497-512)
    e0 = st.PermClsTorchSymmetricTensor(rank=1, dim=2, data={(1,): np.array([1.0, 0.0])}, device=DEV)
    e1 = st.PermClsTorchSymmetricTensor(rank=1, dim=2, data={(1,): np.array([0.0, 1.0])}, device=DEV)
    C = st.multiply.outer(e0, e1)
    assert np.allclose(C["ij"].cpu().numpy(), [0.5]) and np.allclose(C["ii"].cpu().numpy(), [0.0, 0.0])
    # different dimensions: TypeError like the reference (NotImplemented from every provider)
    with pytest.raises(TypeError):
        st.multiply.outer(e0, st.PermClsTorchSymmetricTensor(rank=1, dim=3, data=1.0, device=DEV))


@pytest.mark.parametrize("ra,rb,dim,dist", [(2, 2, 12, "pos"), (3, 2, 9, "normal"), (4, 4, 5, "pos"), (1, 5, 6, "pos"), (3, 3, 10, "normal")])
def test_outer_against_packed_oracle(ra, rb, dim, dist):
    rng = np.random.default_rng(ra * 10 + rb + dim)
    A, B = rand_packed(ra, dim, rng, dist), rand_packed(rb, dim, rng, dist)
    TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=A, device=DEV)
    TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=B, device=DEV)
    C = st.multiply.outer(TA, TB)
    ref, scale = po.outer(A, ra, B, rb, dim), po.outer(absd(A), ra, absd(B), rb, dim)
    assert_classes_close(C, ref, scale, RTOL64)
    # fp32 against the fp64 oracle on the up-cast inputs
    C32 = st.multiply.outer(TA.astype(np.float32), TB.astype(np.float32))
    assert C32.dtype == np.float32
    A32, B32 = {k: v.astype(np.float32).astype(np.float64) for k, v in A.items()}, {k: v.astype(np.float32).astype(np.float64) for k, v in B.items()}
    assert_classes_close(C32, po.outer(A32, ra, B32, rb, dim), scale, RTOL32)
    # identity (A (x)_s B) . x^n = (A . x^ra)(B . x^rb), and the fused outer -> vector kernel
    x = rng.uniform(0.5, 1.5, dim) / np.sqrt(dim)
    lhs = float(st.contract_all_indices_with_vector(C, x))
    rhs = po.contract_all_indices_with_vector(A, ra, dim, x) * po.contract_all_indices_with_vector(B, rb, dim, x)
    s = po.contract_all_indices_with_vector(absd(A), ra, dim, np.abs(x)) * po.contract_all_indices_with_vector(absd(B), rb, dim, np.abs(x))
    assert abs(lhs - rhs) <= 1e-11 * s
    fused = float(ops.outer_then_contract_vec(TA, TB, x))
    assert abs(fused - rhs) <= 1e-11 * s


def test_outer_output_ranges_shard(goldens):
    """[begin, end) output ranges (the multi-GPU partition of multiply.outer): the pieces concatenate to the whole."""
    ra, rb, dim = 3, 3, 8
    rng = np.random.default_rng(5)
    TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=rand_packed(ra, dim, rng), device=DEV)
    TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=rand_packed(rb, dim, rng), device=DEV)
    whole = st.multiply.outer(TA, TB).packed
    total = whole.numel()
    cuts = [min(total, (total * i // 3 + 31) // 32 * 32) for i in range(3)] + [total]
    for b, e in zip(cuts[:-1], cuts[1:]):
        part = torch.empty(e - b, dtype=torch.float64, device=DEV)
        ops.outer_device(TA, TB, part, b, e, torch.float64)
        assert torch.equal(part, whole[b:e])


def test_tensordot_reference_goldens(goldens):
    for c in goldens.cases("tensordot"):
        A, B, ref = goldens.packed(c["tag"] + ".A"), goldens.packed(c["tag"] + ".B"), goldens.packed(c["tag"] + ".out")
        ra, rb, d, k = c["ra"], c["rb"], c["dim"], c["k"]
        TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=d, data=dict(A), device=DEV)
        TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=d, data=dict(B), device=DEV)
        C = st.tensordot(TA, TB, axes=k)
        assert isinstance(C, st.PermClsTorchSymmetricTensor) and C.rank == c["out_rank"] and C.dim == c["out_dim"]
        scale, _ = po.tensordot(absd(full(A, ra, d)), ra, absd(full(B, rb, d)), rb, d, k)
        if C.rank == 0:
            assert abs(float(C) - float(ref[()])) <= 1e-11 * float(scale[()])
            continue
        assert_classes_close(C, ref, scale, 1e-11)


@pytest.mark.parametrize("ra,rb,k,dim", [(3, 3, 1, 11), (3, 3, 2, 9), (4, 2, 1, 8), (3, 1, 1, 17), (4, 3, 2, 6), (2, 2, 2, 13), (3, 2, 1, 40)])
def test_tensordot_against_packed_oracle(ra, rb, k, dim):
    rng = np.random.default_rng(ra + 7 * rb + 31 * k + dim)
    A, B = rand_packed(ra, dim, rng, "normal"), rand_packed(rb, dim, rng, "normal")
    TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=A, device=DEV)
    TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=B, device=DEV)
    ref, n = po.tensordot(A, ra, B, rb, dim, k)
    scale, _ = po.tensordot(absd(A), ra, absd(B), rb, dim, k)
    for axes in (k, (list(range(k)), list(range(k)))):  # NumPy-style axis tuples: only their number matters
        C = st.tensordot(TA, TB, axes=axes)
        assert C.rank == n
        if n == 0:
            assert abs(float(C) - float(ref[()])) <= RTOL64 * float(scale[()])
        else:
            assert_classes_close(C, ref, scale, RTOL64)
    C32 = st.tensordot(TA.astype(np.float32), TB.astype(np.float32), axes=k)
    A32, B32 = {q: v.astype(np.float32).astype(np.float64) for q, v in A.items()}, {q: v.astype(np.float32).astype(np.float64) for q, v in B.items()}
    ref32, _ = po.tensordot(A32, ra, B32, rb, dim, k)
    if n == 0:
        assert abs(float(C32) - float(ref32[()])) <= RTOL32 * float(scale[()])
    else:
        assert_classes_close(C32, ref32, scale, RTOL32)


def test_tensordot_identities():
    """tensordot(A, B, 0) == multiply.outer(A, B) (symtensor/testing/api.py:521-522); r-fold tensordot with a
    vector == contract_all_indices_with_vector."""
    rng = np.random.default_rng(11)
    d = 7
    A, B = rand_packed(3, d, rng), rand_packed(2, d, rng)
    TA = st.PermClsTorchSymmetricTensor(rank=3, dim=d, data=A, device=DEV)
    TB = st.PermClsTorchSymmetricTensor(rank=2, dim=d, data=B, device=DEV)
    assert torch.equal(st.tensordot(TA, TB, axes=0).packed, st.multiply.outer(TA, TB).packed)
    x = rng.uniform(0.5, 1.5, d)
    t = TA
    for _ in range(3):
        t = st.tensordot(t, x, axes=1)
    assert t.rank == 0
    ref = po.contract_all_indices_with_vector(A, 3, d, x)
    assert abs(float(t) - ref) <= 1e-12 * abs(ref)
    assert abs(float(st.contract_all_indices_with_vector(TA, x)) - ref) <= 1e-12 * abs(ref)


def test_matrix_reference_goldens(goldens):
    for c in goldens.cases("mat"):
        A, W, ref = goldens.packed(c["tag"] + ".A"), goldens.ops[c["tag"] + ".W"], goldens.packed(c["tag"] + ".out")
        r, d = c["rank"], c["dim"]
        TA = st.PermClsTorchSymmetricTensor(rank=r, dim=d, data=dict(A), device=DEV)
        C = st.contract_all_indices_with_matrix(TA, W)
        assert isinstance(C, st.PermClsTorchSymmetricTensor) and C.rank == r and C.dim == d
        scale = po.contract_all_indices_with_matrix(absd(full(A, r, d)), r, d, np.abs(W))
        assert_classes_close(C, ref, scale, 1e-11)


@pytest.mark.parametrize("rank,dim", [(2, 33), (3, 20), (4, 12), (5, 7), (6, 6), (3, 70), (4, 66)])
def test_matrix_against_packed_oracle(rank, dim):
    rng = np.random.default_rng(rank * 1000 + dim)
    A = rand_packed(rank, dim, rng, "normal")
    W = rng.standard_normal((dim, dim)) / np.sqrt(dim)
    TA = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=A, device=DEV)
    C = st.contract_all_indices_with_matrix(TA, W)
    ref = po.contract_all_indices_with_matrix(A, rank, dim, W)
    scale = po.contract_all_indices_with_matrix(absd(A), rank, dim, np.abs(W))
    assert_classes_close(C, ref, scale, RTOL64)
    C32 = st.contract_all_indices_with_matrix(TA.astype(np.float32), W.astype(np.float32))
    assert C32.dtype == np.float32
    A32 = {q: v.astype(np.float32).astype(np.float64) for q, v in A.items()}
    ref32 = po.contract_all_indices_with_matrix(A32, rank, dim, W.astype(np.float32).astype(np.float64))
    assert_classes_close(C32, ref32, scale, RTOL32)
    # the flat format goes through the same kernels
    F = st.FlatSymmetricTensor(rank, dim, po.permcls_to_flat(A, rank, dim), device=DEV)
    CF = st.contract_all_indices_with_matrix(F, W)
    assert isinstance(CF, st.FlatSymmetricTensor)
    assert np.allclose(CF.packed.cpu().numpy(), po.permcls_to_flat(ref, rank, dim), rtol=0, atol=1e-9 * max(np.max(v) for v in scale.values()))
    with pytest.raises(ValueError):
        st.contract_all_indices_with_matrix(TA, np.ones((dim + 1, dim)))


def test_matrix_identity_at_config4_shape():
    """BASELINE config 4's shape family (rank 6): vec(mat(A, W), y) == vec(A, W y), beyond the dense oracle's reach
    (rank 6 dim 24: 475,020 packed components, dense would be 1.9e8)."""
    rank, dim = 6, 24
    rng = np.random.default_rng(4)
    t = comb.class_table(rank, dim)
    buf = torch.rand(t.total, dtype=torch.float64, device=DEV) + 0.5
    for c, s, o in zip(t.classes, t.sizes, t.offsets):
        buf[o + s:t.offsets[t.index(c) + 1]] = 0
    A = st.PermClsTorchSymmetricTensor.from_packed(rank, dim, buf)
    W = rng.uniform(0.5, 1.5, (dim, dim)) / dim
    y = rng.uniform(0.5, 1.5, dim)
    C = st.contract_all_indices_with_matrix(A, W)
    lhs = float(st.contract_all_indices_with_vector(C, y))
    rhs = float(st.contract_all_indices_with_vector(A, W @ y))
    assert abs(lhs - rhs) <= 1e-11 * abs(rhs)
    # W = identity returns the tensor itself, bit for bit (index maps of the whole chain)
    I = st.contract_all_indices_with_matrix(A, np.eye(dim))
    assert torch.equal(I.packed, A.packed)


def test_matrix_tensor_pipe_and_register_tile_paths_agree():
    """fp64 mode chain: the FP64-tensor-pipe kernels (gather maps, row strips, last-step kernel) against the DFMA
    register-tile kernel on the same inputs -- same index maps, results to rounding."""
    from symtensor_b200._cabi import c_i64, check, lib
    for rank, dim in [(4, 70), (5, 20), (6, 9), (3, 130)]:
        rng = np.random.default_rng(rank * 100 + dim)
        A = rand_packed(rank, dim, rng, "pos")
        W = rng.uniform(0.5, 1.5, (dim, dim)) / dim
        TA = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=A, device=DEV)
        try:
            check(lib.st_set_tuning(b"mat_dmma", c_i64(0)))
            C0 = st.contract_all_indices_with_matrix(TA, W).packed.clone()
        finally:
            check(lib.st_set_tuning(b"mat_dmma", c_i64(1)))
        C1 = st.contract_all_indices_with_matrix(TA, W).packed
        assert torch.allclose(C0, C1, rtol=1e-13, atol=0)


def test_outer_compile_time_rank_and_run_time_rank_kernels_agree():
    """multiply.outer: the compile-time-rank kernels (shared-memory rank terms, unrolled subsets) against the
    run-time-rank kernel on the same inputs, both operand orders, and the fused outer->vector path."""
    from symtensor_b200 import ops
    from symtensor_b200._cabi import c_i64, check, lib
    for ra, rb, dim in [(4, 4, 7), (2, 4, 9), (3, 3, 11), (3, 1, 40), (1, 1, 50), (4, 3, 6), (2, 2, 70)]:
        rng = np.random.default_rng(ra * 100 + rb * 10 + dim)
        A, B = rand_packed(ra, dim, rng, "pos"), rand_packed(rb, dim, rng, "pos")
        x = rng.uniform(0.5, 1.5, dim)
        TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=A, device=DEV)
        TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=B, device=DEV)
        try:
            check(lib.st_set_tuning(b"outer_fast", c_i64(0)))
            C0 = st.multiply.outer(TA, TB).packed.clone()
            v0 = float(ops.outer_then_contract_vec(TA, TB, x))
        finally:
            check(lib.st_set_tuning(b"outer_fast", c_i64(1)))
        C1 = st.multiply.outer(TA, TB).packed
        v1 = float(ops.outer_then_contract_vec(TA, TB, x))
        assert torch.allclose(C0, C1, rtol=1e-13, atol=0), (ra, rb, dim)
        assert abs(v0 - v1) <= 1e-12 * abs(v0)


def test_row_walk_on_the_device_equals_its_host_replay():
    """The warp-uniform row walk as the kernels run it (rowwalk_debug_kernel: seek per span, batches of 32 lanes, latch, lane-side
    decode) against the host replay of the same code (st_debug_rowwalk, which tests/test_oracle_index.py pins to the oracle's
    enumeration): every coordinate, several span lengths.  (A latch left undefined for unserved lanes once compiled into a
    kernel whose lanes all saw the last row of the call -- the host replay could not see that.)"""
    from symtensor_b200 import combinatorics as comb
    from symtensor_b200._cabi import c_i64, check, lib
    for rank, dim, span in [(1, 7, 2048), (2, 9, 2048), (3, 6, 64), (4, 11, 2048), (5, 5, 96), (6, 7, 2048), (8, 5, 2048), (8, 9, 2048),
                            (4, 40, 1024), (3, 255, 2048), (7, 6, 32), (8, 12, 4096)]:
        total = comb.class_table(rank, dim).total
        host = np.full((total, rank), -7, dtype=np.int32)
        assert lib.st_debug_rowwalk(rank, c_i64(dim), c_i64(0), c_i64(total), c_i64(span), host.ctypes.data) == total
        dev = torch.full((total, rank), -9, dtype=torch.int32, device=DEV)
        check(lib.st_debug_rowwalk_device(rank, c_i64(dim), c_i64(0), c_i64(total), c_i64(span), dev.data_ptr(), None))
        torch.cuda.synchronize()
        got = dev.cpu().numpy()
        bad = np.nonzero((got != host).any(axis=1))[0]
        assert len(bad) == 0, (rank, dim, span, bad[:5], got[bad[:3]], host[bad[:3]])


def test_row_walk_layout_converters_agree_with_the_unrank_kernels():
    """permcls <-> flat re-ordering through the row-walk kernels (rowwalk_convert_kernel) against the kernels that unrank /
    classify / rank every component: bit-identical buffers, including output ranges that start and end mid-row."""
    from symtensor_b200 import combinatorics as comb
    from symtensor_b200._cabi import c_i64, check, lib
    for rank, dim in [(1, 9), (2, 40), (3, 37), (4, 12), (5, 9), (6, 8), (8, 5), (8, 3), (4, 255), (7, 6)]:
        tab = comb.class_table(rank, dim)
        n_flat = comb.indep_size(rank, dim)
        for tdt, fn in ((torch.float64, "f64"), (torch.float32, "f32")):
            src = torch.rand(tab.total, dtype=tdt, device=DEV)
            got = {}
            for rows in (0, 1):
                try:
                    check(lib.st_set_tuning(b"conv_rows", c_i64(rows)))
                    flat = torch.full((n_flat,), -5.0, dtype=tdt, device=DEV)
                    check(getattr(lib, "st_permcls_to_flat_" + fn)(rank, c_i64(dim), src.data_ptr(), flat.data_ptr(), None))
                    b, e = tab.total // 3 + 3, tab.total - 2
                    back = torch.full((tab.total,), -5.0, dtype=tdt, device=DEV)
                    part = torch.full((e - b,), -5.0, dtype=tdt, device=DEV)
                    check(getattr(lib, "st_flat_to_permcls_" + fn)(rank, c_i64(dim), flat.data_ptr(), back.data_ptr(), c_i64(0), c_i64(tab.total), None))
                    check(getattr(lib, "st_flat_to_permcls_" + fn)(rank, c_i64(dim), flat.data_ptr(), part.data_ptr(), c_i64(b), c_i64(e), None))
                    torch.cuda.synchronize()
                    got[rows] = (flat, back, part)
                finally:
                    check(lib.st_set_tuning(b"conv_rows", c_i64(1)))
            for x0, x1 in zip(got[0], got[1]):
                assert torch.equal(x0, x1), (rank, dim, fn)
            assert float(got[1][0].min()) >= 0.0  # every flat position was written


def test_matrix_pipeline_kernel_agrees_with_the_first_dmma_kernel():
    """contract_all_indices_with_matrix (fp64, dim <= 64): the persistent producer / consumer step kernel (st_mat.cu) against
    the per-CTA kernel of round 1, whole tensors and first-mode partitions."""
    from symtensor_b200._cabi import c_i64, check, lib
    for rank, dim in [(2, 64), (3, 33), (4, 20), (5, 11), (6, 9), (3, 64), (4, 7)]:
        rng = np.random.default_rng(rank * 100 + dim)
        TA = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=rand_packed(rank, dim, rng, "normal"), device=DEV)
        W = rng.standard_normal((dim, dim))
        try:
            check(lib.st_set_tuning(b"mat_pipe", c_i64(0)))
            C0 = st.contract_all_indices_with_matrix(TA, W).packed.clone()
        finally:
            check(lib.st_set_tuning(b"mat_pipe", c_i64(1)))
        C1 = st.contract_all_indices_with_matrix(TA, W).packed.clone()
        assert torch.allclose(C0, C1, rtol=1e-13, atol=1e-13), (rank, dim)
        # gather positions ranked by the producers (default: only steps whose map does not fit) against the gather maps
        for rows in (1, 1 << 40):
            try:
                check(lib.st_set_tuning(b"mat_onfly_rows", c_i64(rows)))
                C2 = st.contract_all_indices_with_matrix(TA, W).packed
                assert torch.equal(C1, C2), (rank, dim, rows)
            finally:
                check(lib.st_set_tuning(b"mat_onfly_rows", c_i64(0)))


def test_tensordot_several_output_ranges_in_one_call_equal_the_single_range_calls():
    """st_tensordot_ranges_f32 (the multi-GPU shard: up to 8 disjoint ranges, a tile that serves several of them runs once)
    against one st_tensordot_f32 call per range, through the tiled tcgen05 kernel (dims below its default threshold are
    admitted with the tuning key) and through the materialised path: bit-identical."""
    from symtensor_b200 import combinatorics as comb, ops, sharding
    from symtensor_b200._cabi import c_i64, check, lib
    for ra, rb, k, dim, tiled in [(3, 3, 1, 40, True), (3, 3, 1, 24, True), (4, 4, 2, 9, True), (3, 2, 1, 30, False)]:
        rng = np.random.default_rng(ra * 1000 + rb * 100 + k * 10 + dim)
        TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=rand_packed(ra, dim, rng, "normal"), device=DEV).astype(np.float32)
        TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=rand_packed(rb, dim, rng, "normal"), device=DEV).astype(np.float32)
        n = ra + rb - 2 * k
        total = comb.class_table(n, dim).total
        try:
            check(lib.st_set_tuning(b"sym22_min_dim", c_i64(8 if tiled else 96)))
            assert bool(lib.st_tensordot_is_tiled(ra, rb, k, c_i64(dim), 4)) == tiled
            if n == 4:
                sets = sharding.tensordot22_shards(dim, 3)
            else:
                cuts = sharding.shard_bounds(total, 5)
                sets = [[(cuts[0], cuts[1]), (cuts[3], cuts[4])], [(cuts[1], cuts[3]), (cuts[4], cuts[5])]]
            for ranges in sets:
                outs = [torch.full((e - b,), -3.0, dtype=torch.float32, device=DEV) for b, e in ranges]
                ops.tensordot_device_ranges(TA, TB, k, outs, ranges)
                for (b, e), got in zip(ranges, outs):
                    one = torch.full((e - b,), -5.0, dtype=torch.float32, device=DEV)
                    ops.tensordot_device(TA, TB, k, one, b, e, torch.float32)
                    assert torch.equal(got, one), (ra, rb, k, dim, b, e)
                if tiled:  # several batches of tiles per call (BASELINE config 3 runs 7 batches of 32768 per GPU)
                    check(lib.st_set_tuning(b"sym22_batch_tiles", c_i64(3)))
                    again = [torch.full((e - b,), -9.0, dtype=torch.float32, device=DEV) for b, e in ranges]
                    ops.tensordot_device_ranges(TA, TB, k, again, ranges)
                    check(lib.st_set_tuning(b"sym22_batch_tiles", c_i64(32768)))
                    for got, ref in zip(again, outs):
                        assert torch.equal(got, ref), (ra, rb, k, dim)
        finally:
            check(lib.st_set_tuning(b"sym22_min_dim", c_i64(96)))
            check(lib.st_set_tuning(b"sym22_batch_tiles", c_i64(32768)))


def test_tensordot_ranges_argument_checks():
    """st_tensordot_ranges_f32 rejects overlapping ranges, ranges outside the packed output and more than 8 ranges (status
    ST_ERR_INVALID -> ValueError), and accepts empty ranges."""
    from symtensor_b200 import combinatorics as comb, ops
    from symtensor_b200._cabi import c_i64, check, lib
    dim = 24
    rng = np.random.default_rng(1)
    TA = st.PermClsTorchSymmetricTensor(rank=3, dim=dim, data=rand_packed(3, dim, rng, "normal"), device=DEV).astype(np.float32)
    total = comb.class_table(4, dim).total
    try:
        check(lib.st_set_tuning(b"sym22_min_dim", c_i64(8)))
        buf = lambda n: torch.zeros(max(n, 1), dtype=torch.float32, device=DEV)
        with pytest.raises(ValueError, match="overlap"):
            ops.tensordot_device_ranges(TA, TA, 1, [buf(64), buf(64)], [(0, 64), (32, 96)])
        with pytest.raises(ValueError):
            ops.tensordot_device_ranges(TA, TA, 1, [buf(64)], [(total - 32, total + 32)])
        with pytest.raises(ValueError):
            ops.tensordot_device_ranges(TA, TA, 1, [buf(32)] * 9, [(32 * i, 32 * i + 32) for i in range(9)])
        one, two = buf(64), buf(0)
        ops.tensordot_device_ranges(TA, TA, 1, [one, two], [(0, 64), (128, 128)])  # an empty range is fine
        ref = buf(64)
        ops.tensordot_device(TA, TA, 1, ref, 0, 64, torch.float32)
        assert torch.equal(one, ref)
    finally:
        check(lib.st_set_tuning(b"sym22_min_dim", c_i64(96)))


def test_outer_row_walk_kernel_agrees_with_the_per_component_unrank_kernels():
    """multiply.outer through the row-walk kernel (outer_rows_kernel: warp-uniform odometer over the rows that cover 32
    consecutive coordinates, half-split subset sums) against the kernel that unranks every component (outer_fast_kernel) and
    the run-time-rank kernel: whole tensors, ranges that begin and end mid-row / mid-class, fp64 and fp32, fused vector path."""
    from symtensor_b200 import combinatorics as comb, ops
    from symtensor_b200._cabi import c_i64, check, lib
    for ra, rb, dim in [(4, 4, 7), (4, 4, 3), (2, 4, 9), (3, 3, 11), (3, 1, 40), (1, 1, 50), (4, 3, 6), (2, 2, 70), (1, 4, 12),
                        (2, 1, 255), (4, 4, 10)]:
        rng = np.random.default_rng(ra * 100 + rb * 10 + dim)
        A, B = rand_packed(ra, dim, rng, "pos"), rand_packed(rb, dim, rng, "pos")
        x = rng.uniform(0.5, 1.5, dim)
        TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=A, device=DEV)
        TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=B, device=DEV)
        total = comb.class_table(ra + rb, dim).total
        ranges = [(0, total), (total // 3 + 5, total - 7), (max(0, total // 2 - 1), min(total, total // 2 + 4099))]
        for tdt in (torch.float64, torch.float32):
            TAt, TBt = (TA, TB) if tdt == torch.float64 else (TA.astype(np.float32), TB.astype(np.float32))
            outs = {}
            for rows in (0, 1):
                try:
                    check(lib.st_set_tuning(b"outer_rows", c_i64(rows)))
                    got = []
                    for b, e in ranges:
                        part = torch.full((e - b,), -777.0, dtype=tdt, device=DEV)
                        ops.outer_device(TAt, TBt, part, b, e, tdt)
                        got.append(part)
                    outs[rows] = (got, float(ops.outer_then_contract_vec(TAt, TBt, x.astype(np.float32 if tdt == torch.float32 else np.float64))))
                finally:
                    check(lib.st_set_tuning(b"outer_rows", c_i64(1)))
            for g0, g1 in zip(outs[0][0], outs[1][0]):
                if tdt == torch.float64:
                    assert torch.allclose(g0, g1, rtol=1e-13, atol=1e-300), (ra, rb, dim)
                else:  # fp32 (positive data): the two kernels round the sum of the products differently
                    assert torch.allclose(g0, g1, rtol=2e-6, atol=0), (ra, rb, dim)
                assert torch.equal(g0 == 0, g1 == 0)  # the alignment padding
            v0, v1 = outs[0][1], outs[1][1]
            assert abs(v0 - v1) <= (1e-11 if tdt == torch.float64 else 2e-4) * max(abs(v0), 1e-3), (ra, rb, dim, v0, v1)


def test_tensordot_fp32_tensor_core_gram_matches_cuda_core_gram_and_oracle():
    """fp32 tensordot: the tcgen05 (3xTF32, two-level accumulation) Gram kernel against the CUDA-core kernel and the
    fp64 oracle, including contraction lengths beyond one accumulation chain (K > 256) and K not a multiple of 4."""
    from symtensor_b200._cabi import c_i64, check, lib
    for ra, rb, k, dim in [(2, 2, 1, 600), (3, 2, 1, 70), (2, 2, 1, 301), (3, 3, 2, 24), (2, 1, 1, 1000)]:
        rng = np.random.default_rng(ra * 1000 + rb * 100 + k * 10 + dim)
        A, B = rand_packed(ra, dim, rng, "normal"), rand_packed(rb, dim, rng, "normal")
        TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=A, device=DEV).astype(np.float32)
        TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=B, device=DEV).astype(np.float32)
        A32 = {q: v.astype(np.float32).astype(np.float64) for q, v in A.items()}
        B32 = {q: v.astype(np.float32).astype(np.float64) for q, v in B.items()}
        ref, n = po.tensordot(A32, ra, B32, rb, dim, k)
        scale, _ = po.tensordot(absd(A32), ra, absd(B32), rb, dim, k)
        try:  # the CUDA-core Gram kernel and the run-time-rank epilogue
            check(lib.st_set_tuning(b"gram_umma", c_i64(0)))
            check(lib.st_set_tuning(b"outer_fast", c_i64(0)))
            C0 = st.tensordot(TA, TB, axes=k)
            assert_classes_close(C0, ref, scale, RTOL32)
        finally:
            check(lib.st_set_tuning(b"gram_umma", c_i64(1)))
            check(lib.st_set_tuning(b"outer_fast", c_i64(1)))
        C1 = st.tensordot(TA, TB, axes=k)
        assert_classes_close(C1, ref, scale, RTOL32)


@pytest.mark.parametrize("kch", [16, 32])
@pytest.mark.parametrize("ra,rb,k,dim", [(3, 3, 1, 11), (3, 3, 1, 24), (3, 3, 1, 40), (4, 4, 2, 9), (3, 3, 1, 70)])
def test_tensordot_tiled_tcgen05_kernel_against_packed_oracle(ra, rb, k, dim, kch):
    """fp32 tensordot with two free indices on each side (BASELINE config 3's shape family) through the NON-MATERIALISING
    kernel (st_sym22.cu: TMA-staged operand boxes, three 128 x 256 tcgen05 GEMMs per 8 x 16 x 16 x 16 output tile, chains of
    256 added in registers, red.global.add into the packed output): the whole output against the fp64 packed oracle on the
    up-cast inputs (1e-5 of the component's sum of |terms|), both stage geometries, and [begin, end) output ranges -- the
    multi-GPU partition -- concatenating bit for bit to the whole."""
    from symtensor_b200._cabi import c_i64, check, lib
    rng = np.random.default_rng(ra * 1000 + rb * 100 + k * 10 + dim)
    dist = "pos" if dim % 2 else "normal"
    A, B = rand_packed(ra, dim, rng, dist), rand_packed(rb, dim, rng, dist)
    TA = st.PermClsTorchSymmetricTensor(rank=ra, dim=dim, data=A, device=DEV).astype(np.float32)
    TB = st.PermClsTorchSymmetricTensor(rank=rb, dim=dim, data=B, device=DEV).astype(np.float32)
    A32 = {q: v.astype(np.float32).astype(np.float64) for q, v in A.items()}
    B32 = {q: v.astype(np.float32).astype(np.float64) for q, v in B.items()}
    ref, n = po.tensordot(A32, ra, B32, rb, dim, k)
    scale, _ = po.tensordot(absd(A32), ra, absd(B32), rb, dim, k)
    try:
        check(lib.st_set_tuning(b"sym22_min_dim", c_i64(1)))
        check(lib.st_set_tuning(b"sym22_kch", c_i64(kch)))
        assert lib.st_tensordot_is_tiled(ra, rb, k, c_i64(dim), 4) == 1 and lib.st_tensordot_is_tiled(ra, rb, k, c_i64(dim), 8) == 0
        C = st.tensordot(TA, TB, axes=k)
        assert n == 4 and C.rank == 4 and C.dtype == np.float32
        assert_classes_close(C, ref, scale, RTOL32)
        whole = C.packed
        total = whole.numel()
        cuts = sorted({0, total} | {min(total, int(total * f) // 32 * 32) for f in (0.013, 0.21, 0.5, 0.77)})
        for b, e in zip(cuts[:-1], cuts[1:]):
            part = torch.full((e - b,), float("nan"), dtype=torch.float32, device=DEV)
            ops.tensordot_device(TA, TB, k, part, b, e, torch.float32)
            assert torch.equal(part, whole[b:e]), (b, e)
    finally:
        check(lib.st_set_tuning(b"sym22_min_dim", c_i64(96)))
        check(lib.st_set_tuning(b"sym22_kch", c_i64(32)))


@pytest.mark.parametrize("rank,dim,dtype", [(4, 12, np.float64), (3, 20, np.float64), (6, 6, np.float64), (5, 9, np.float32), (2, 33, np.float64), (4, 70, np.float64)])
def test_matrix_first_mode_ranges_concatenate_to_the_whole(rank, dim, dtype):
    """The multi-GPU partition of contract_all_indices_with_matrix by the FIRST output mode (st_contract_mat_range_*): the slices
    of 1, 2, 3 and 8 GPUs -- each computed from its own slice of the mode chain -- concatenate bit for bit to the flat-ordered
    result of the whole contraction, and that agrees with the packed oracle (symtensor/symalg.py:475-496)."""
    from symtensor_b200 import sharding
    rng = np.random.default_rng(rank * 37 + dim)
    A = rand_packed(rank, dim, rng, "normal")
    W = (rng.standard_normal((dim, dim)) / np.sqrt(dim)).astype(dtype)
    TA = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=A, device=DEV).astype(dtype)
    whole = ops._flat_buffer(st.contract_all_indices_with_matrix(TA, W), TA.torch_dtype)
    if dim <= 20:
        A_in = {q: v.astype(dtype).astype(np.float64) for q, v in A.items()}
        ref = po.permcls_to_flat(po.contract_all_indices_with_matrix(A_in, rank, dim, W.astype(np.float64)), rank, dim)
        scale = po.permcls_to_flat(po.contract_all_indices_with_matrix(absd(A_in), rank, dim, np.abs(W).astype(np.float64)), rank, dim)
        tol = RTOL64 if dtype == np.float64 else RTOL32
        assert np.all(np.abs(whole.cpu().numpy().astype(np.float64) - ref) <= tol * np.maximum(scale, 1e-300))
    for world in (1, 2, 3, 8):
        cuts = sharding.mat_mode_bounds(rank, dim, world)
        assert cuts[0] == 0 and cuts[-1] == dim and all(a <= b for a, b in zip(cuts[:-1], cuts[1:]))
        pos = 0
        for jlo, jhi in zip(cuts[:-1], cuts[1:]):
            part, b, e = ops.contract_mat_device(TA, W, jlo, jhi)
            assert b == pos and e - b == part.numel()
            assert torch.equal(part, whole[b:e]), (world, jlo, jhi)
            pos = e
        assert pos == whole.numel()
