"""Dense <-> packed on the GPU (pack_dense_kernel / unpack_dense_kernel, csrc/st_pack.cu): the constructor-from-dense and
``todense`` steps either side of the ops (symtensor/permcls_symtensor.py:599-618, 883-887; symtensor/flat_symtensor.py:100-110,
251-256), against the dense oracle (which restates the reference's loops and is pinned to its golden vectors).

Bit-exact: packing and unpacking move values, they do no arithmetic (symmetrize averages up to rank! terms: 1e-13 of the mean of |terms|)."""
import itertools

import numpy as np
import pytest
import torch

from oracle import dense_oracle as do
from oracle import index_oracle as io
from oracle import packed_oracle as po

import symtensor_b200 as st
from symtensor_b200._cabi import LAYOUT_FLAT, LAYOUT_PERMCLS, c_i64, check, lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(1, 7), (2, 9), (3, 6), (3, 2), (4, 11), (5, 5), (6, 4), (8, 3), (4, 1)]


def rand_packed(rank, dim, rng, dtype=np.float64):
    return {c: rng.standard_normal(io.permclass_size(c, dim)).astype(dtype) for c in io.perm_classes(rank)}


@pytest.mark.parametrize("rank,dim", SHAPES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_todense_matches_the_dense_oracle_bit_exactly(rank, dim, dtype):
    rng = np.random.default_rng(1000 * rank + dim)
    data = rand_packed(rank, dim, rng, dtype)
    ref = do.todense(data, rank, dim, dtype=dtype)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device=DEV)
    got = A.todense()
    assert got.dtype == (torch.float64 if dtype == np.float64 else torch.float32) and tuple(got.shape) == (dim,) * rank
    assert np.array_equal(got.cpu().numpy(), ref)
    F = st.FlatSymmetricTensor(rank, dim, data=po.permcls_to_flat(data, rank, dim, dtype=dtype), device=DEV)
    assert np.array_equal(F.todense().cpu().numpy(), ref)


@pytest.mark.parametrize("rank,dim", SHAPES)
def test_construction_from_a_symmetric_dense_array_round_trips(rank, dim):
    rng = np.random.default_rng(2000 * rank + dim)
    data = rand_packed(rank, dim, rng)
    dense = do.todense(data, rank, dim)
    ref = do.repack(dense, rank, dim)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=dense, device=DEV)  # numpy input: H2D, then the pack kernel
    got = A.to_numpy_dict()
    for c, r in ref.items():
        assert np.array_equal(np.asarray(got[c]), r), c
    assert float(A.packed.abs().sum()) == pytest.approx(sum(float(np.abs(v).sum()) for v in ref.values()), rel=1e-12)  # zero padding
    B = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=torch.as_tensor(dense, device=DEV), device=DEV)
    assert torch.equal(B.packed, A.packed)
    F = st.FlatSymmetricTensor(rank, dim, data=dense, device=DEV)
    assert np.array_equal(F.packed.cpu().numpy(), po.permcls_to_flat(ref, rank, dim))
    # A[:] = dense (symtensor/permcls_symtensor.py:809-819)
    C = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, device=DEV)
    C[:] = dense
    assert torch.equal(C.packed, A.packed)


@pytest.mark.parametrize("rank,dim", [(2, 9), (3, 6), (4, 7), (5, 4), (6, 3)])
def test_symmetrize_is_the_mean_over_all_axis_permutations(rank, dim):
    rng = np.random.default_rng(3000 * rank + dim)
    dense = rng.standard_normal((dim,) * rank)
    ref = do.repack(do.symmetrize(dense), rank, dim)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=dense, symmetrize=True, device=DEV)
    got = A.to_numpy_dict()
    scale = do.repack(do.symmetrize(np.abs(dense)), rank, dim)
    for c, r in ref.items():
        assert np.all(np.abs(np.asarray(got[c]) - r) <= 1e-13 * scale[c] + 1e-300), c
    F = st.FlatSymmetricTensor(rank, dim, data=dense, symmetrize=True, device=DEV)
    assert np.all(np.abs(F.packed.cpu().numpy() - po.permcls_to_flat(ref, rank, dim)) <= 1e-13 * po.permcls_to_flat(scale, rank, dim) + 1e-300)


def test_asymmetric_dense_arrays_are_rejected_like_the_reference():
    """utils.is_symmetric: np.allclose(rtol=1e-5, atol=1e-8) over all axis permutations (symtensor/utils.py:563-578);
    the constructor raises ValueError("Data array is not symmetric.") (symtensor/permcls_symtensor.py:610-613)."""
    rng = np.random.default_rng(7)
    for rank, dim in [(2, 5), (3, 4), (4, 3)]:
        data = rand_packed(rank, dim, rng)
        dense = do.todense(data, rank, dim)
        idx = tuple(range(1, rank)) + (0,) if dim >= rank else (1,) + (0,) * (rank - 1)
        for delta, ok in [(1e-3, False), (1e-7 * abs(dense[idx]), True), (np.nan, False)]:
            bad = dense.copy()
            bad[idx] += delta  # one non-representative permutation of one component
            ref_ok = all(np.allclose(bad, bad.transpose(p)) for p in itertools.permutations(range(rank)))
            assert ref_ok == ok
            if ok:
                st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=bad, device=DEV)
                st.FlatSymmetricTensor(rank, dim, data=bad, device=DEV)
            else:
                with pytest.raises(ValueError, match="not symmetric"):
                    st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=bad, device=DEV)
                with pytest.raises(RuntimeError, match="not symmetric"):
                    st.FlatSymmetricTensor(rank, dim, data=bad, device=DEV)
    with pytest.raises(ValueError):  # shape that does not broadcast to dim^rank
        st.PermClsTorchSymmetricTensor(rank=3, dim=4, data=np.zeros((4, 4, 3)), device=DEV)


def test_pack_ranges_and_argument_errors_through_the_c_abi():
    rank, dim = 4, 9
    rng = np.random.default_rng(11)
    data = rand_packed(rank, dim, rng)
    dense = torch.as_tensor(do.todense(data, rank, dim), device=DEV)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device=DEV)
    total = A.packed.numel()
    out = torch.full((total,), -1.0, dtype=torch.float64, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    cut = (total // 2) // 32 * 32
    for b, e in [(0, cut), (cut, total)]:  # two shards of the packed range
        check(lib.st_pack_dense_f64(LAYOUT_PERMCLS, rank, c_i64(dim), dense.data_ptr(), out[b:].data_ptr(), c_i64(b), c_i64(e), 0, 1e-5, 1e-8,
                                    flag.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(out, A.packed) and int(flag.item()) == 0
    assert lib.st_pack_dense_f64(LAYOUT_PERMCLS, rank, c_i64(dim), dense.data_ptr(), out.data_ptr(), c_i64(0), c_i64(total + 1), 0, 1e-5, 1e-8,
                                 flag.data_ptr(), None) != 0
    assert lib.st_pack_dense_f64(7, rank, c_i64(dim), dense.data_ptr(), out.data_ptr(), c_i64(0), c_i64(total), 0, 1e-5, 1e-8, flag.data_ptr(), None) != 0
    assert lib.st_unpack_dense_f64(LAYOUT_FLAT, rank, c_i64(dim), None, dense.data_ptr(), None) != 0
    assert lib.st_unpack_dense_f64(LAYOUT_FLAT, 12, c_i64(10 ** 6), out.data_ptr(), dense.data_ptr(), None) != 0  # dim^rank overflows


def test_dense_round_trip_at_a_size_beyond_the_oracle():
    """rank 4 dim 64 (1.7e7 dense elements, 766,480 components): todense -> pack is the identity, and todense is symmetric."""
    rank, dim = 4, 64
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, device=DEV)
    torch.manual_seed(5)
    for v in A.values():
        v.copy_(torch.randn(v.shape, dtype=torch.float64, device=DEV))
    dense = A.todense()
    assert torch.equal(dense, dense.permute(3, 1, 0, 2)) and torch.equal(dense, dense.permute(1, 2, 3, 0))
    B = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=dense, device=DEV)
    assert torch.equal(B.packed, A.packed)


def test_host_resident_tensors_pack_and_unpack_through_the_gpu_kernels():
    """device="host" tensors (the end-to-end path with host buffers) run todense / the constructor from a dense array through
    the same CUDA kernels (copy in, kernel, copy out) -- the index-gather loops in permcls.py / flat.py are only reached on a
    machine without any CUDA device: the launch counter of the library grows and the results equal the device-resident ones."""
    rng = np.random.default_rng(77)
    for rank, dim in [(3, 6), (4, 5), (2, 9)]:
        data = rand_packed(rank, dim, rng)
        ref = do.todense(data, rank, dim)
        n0 = lib.st_launch_count()
        H = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device="host")
        dense = H.todense()
        assert dense.device.type == "cpu" and np.array_equal(dense.numpy(), ref)
        assert lib.st_launch_count() > n0
        n1 = lib.st_launch_count()
        H2 = st.PermClsTorchSymmetricTensor(data=ref, device="host")
        assert lib.st_launch_count() > n1
        assert all(np.array_equal(H2.to_numpy_dict()[c], data[c]) for c in data)
        bad = ref.copy()
        bad[(0,) * (rank - 1) + (1,)] += 1.0
        with pytest.raises(ValueError, match="not symmetric"):
            st.PermClsTorchSymmetricTensor(data=bad, device="host")
        F = st.FlatSymmetricTensor(rank, dim, data=ref, device="host")
        assert np.array_equal(F.todense().numpy(), ref)
