"""The BASELINE configurations at their REAL sizes (BASELINE.json configs[2..4]; configs[0..1] are in test_gpu_vec.py).

The reference cannot run any of them (dense arrays of 4 TB / 550 GB / 26 TB), and the packed NumPy
oracle cannot produce a whole result in seconds either, so parity at full size is established by
  * the packed formulas of ``oracle/packed_oracle.py`` evaluated on RANDOM WINDOWS of the output (component-wise, fp64 on
    the up-cast inputs; tolerance 1e-12 relative (fp64) / 1e-5 relative (fp32) of the component's sum of |terms|),
  * bit-exact index-map checks (identity and permutation matrices for the mode chain: every product is x * 1.0 or x * 0.0,
    so any wrong gather map changes bits), and
  * the size-independent identities of SURVEY.md 8c.

C3  rank 3 dim 1000 fp32 tensordot over one index          symtensor/symalg.py:427-459
C4  rank 6 dim 64 fp64 contract_all_indices_with_matrix     symtensor/symalg.py:475-496
C5  rank 8 dim 40 fp32 multiply.outer of two rank-4 tensors, then the vector contraction   symtensor/symalg.py:295-316, 505-527
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import index_oracle as io
from oracle import packed_oracle as po

import symtensor_b200 as st
from symtensor_b200 import combinatorics as comb
from symtensor_b200 import ops, sharding

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def device_tensor(rank, dim, seed, dtype=torch.float64):
    """Seeded synthetic tensor generated on the device, values U[0.5, 1.5), alignment padding zero."""
    t = comb.class_table(rank, dim)
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    buf = (torch.rand(t.total, generator=g, dtype=torch.float64, device=DEV) + 0.5).to(dtype)
    for i in range(t.ncls):
        buf[t.offsets[i] + t.sizes[i]:t.offsets[i + 1]] = 0
    return st.PermClsTorchSymmetricTensor.from_packed(rank, dim, buf)


def sample_positions(size, n, rng):
    """Window starts + random single positions of a class: the first and last components, and n random ones."""
    if size == 0:
        return np.zeros(0, dtype=np.int64)
    pos = np.concatenate([np.arange(min(size, 8)), np.arange(max(0, size - 8), size), rng.integers(0, size, n)])
    return np.unique(pos).astype(np.int64)


def flat_host(T, dtype=np.float64):
    """The components of a (small or medium) tensor in flat order on the host, up-cast."""
    return ops._flat_buffer(T, T.torch_dtype).cpu().numpy().astype(dtype)


def rep_index(cls, dim, pos):
    """Representative multi-index of component `pos` of class `cls`: the distinct values (class order) repeated by their
    multiplicities (get_index_representative, symtensor/permcls_symtensor.py:375-381)."""
    vals = io.permcls_unrank(cls, dim, int(pos))
    return tuple(v for v, m in zip(vals, cls) for _ in range(m))


def sorted_indices(cls, dim, positions):
    return np.array([sorted(rep_index(cls, dim, p)) for p in positions], dtype=np.int64).reshape(len(positions), sum(cls))


# ------------------------------------------------------------------------------------------------------------
# C4: rank 6 dim 64 fp64 matrix contraction
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c4_tensor():
    return device_tensor(6, 64, 20261018 + 4)


def _gather_components(buf_host, table, dim, indices):
    """A[index] for arbitrary (unsorted) multi-indices through the ORACLE's rank (class, position)."""
    out = np.empty(len(indices))
    for n, idx in enumerate(indices):
        cls, pos = io.rank_of_index(tuple(int(v) for v in idx), dim)
        out[n] = buf_host[table.offsets[table.index(cls)] + pos]
    return out


def test_config4_identity_and_permutation_are_bit_exact(c4_tensor):
    """W = I returns the tensor bit for bit; W = a permutation matrix P (W[p(j), j] = 1) returns C[J] = A[p(J)] bit for bit
    -- every index map of the 6-step chain, at full size (119,877,472 components, intermediates up to 16.7 GB)."""
    A = c4_tensor
    rank, dim = 6, 64
    I = st.contract_all_indices_with_matrix(A, np.eye(dim))
    assert torch.equal(I.packed, A.packed)
    del I
    rng = np.random.default_rng(44)
    p = rng.permutation(dim)
    W = np.zeros((dim, dim))
    W[p, np.arange(dim)] = 1.0
    C = st.contract_all_indices_with_matrix(A, W)
    table = A.class_table
    a_host = A.packed.cpu().numpy()
    c_host = C.packed.cpu().numpy()
    del C
    for cls in table.classes:
        ci = table.index(cls)
        pos = sample_positions(table.sizes[ci], 150, rng)
        J = np.array([rep_index(cls, dim, q) for q in pos], dtype=np.int64)
        want = _gather_components(a_host, table, dim, p[J])
        got = c_host[table.offsets[ci] + pos]
        assert np.array_equal(got, want), cls


def test_config4_windows_against_the_packed_direct_form(c4_tensor):
    """Component-wise parity on random windows of every output class: with two non-zeros per column of W the direct form
    C_J = sum_I A_I prod_k W[i_k, j_k] (symtensor/symalg.py:492-495) has 2^6 terms per component, which the oracle sums
    in fp64 from the host copy of A; plus the dense-W identity vec(mat(A, W), y) == vec(A, W y)."""
    A = c4_tensor
    rank, dim = 6, 64
    rng = np.random.default_rng(46)
    W = np.zeros((dim, dim))
    rows = np.stack([rng.permutation(dim), rng.permutation(dim)])
    rows[1] = np.where(rows[1] == rows[0], (rows[1] + 1) % dim, rows[1])
    for j in range(dim):
        W[rows[0, j], j] = rng.uniform(0.5, 1.5)
        W[rows[1, j], j] = rng.uniform(0.5, 1.5)
    C = st.contract_all_indices_with_matrix(A, W)
    table = A.class_table
    a_host = A.packed.cpu().numpy()
    c_host = C.packed.cpu().numpy()
    del C
    choices = list(itertools.product((0, 1), repeat=rank))
    checked = 0
    for cls in table.classes:
        ci = table.index(cls)
        pos = sample_positions(table.sizes[ci], 40, rng)
        for q in pos:
            J = rep_index(cls, dim, q)
            terms = []
            for ch in choices:
                I = [rows[b, j] for b, j in zip(ch, J)]
                w = np.prod([W[i, j] for i, j in zip(I, J)])
                terms.append(w * _gather_components(a_host, table, dim, [I])[0])
            want = float(np.sum(terms))
            got = c_host[table.offsets[ci] + q]
            assert abs(got - want) <= 1e-12 * float(np.sum(np.abs(terms))), (cls, int(q), got, want)
            checked += 1
    assert checked > 400
    # dense random W: the identity at full size
    Wd = rng.uniform(0.5, 1.5, (dim, dim)) / dim
    y = rng.uniform(0.5, 1.5, dim)
    Cd = st.contract_all_indices_with_matrix(A, Wd)
    lhs = float(st.contract_all_indices_with_vector(Cd, y))
    rhs = float(st.contract_all_indices_with_vector(A, Wd @ y))
    assert abs(lhs - rhs) <= 1e-12 * abs(rhs), (lhs, rhs)


def test_config4_first_mode_partition_of_8_gpus(c4_tensor):
    """BASELINE config 4 cut for 8 GPUs by the first output mode (sharding.mat_mode_bounds): every slice -- computed from its own
    slice of the mode chain, with its own (smaller) workspace -- equals the corresponding range of the whole result bit for bit."""
    from symtensor_b200 import sharding
    A = c4_tensor
    rank, dim = 6, 64
    rng = np.random.default_rng(48)
    W = rng.uniform(0.5, 1.5, (dim, dim)) / dim
    whole = ops._flat_buffer(st.contract_all_indices_with_matrix(A, W), torch.float64)
    cuts = sharding.mat_mode_bounds(rank, dim, 8)
    af = ops._flat_buffer(A, torch.float64)
    pos = 0
    for jlo, jhi in zip(cuts[:-1], cuts[1:]):
        part, b, e = ops.contract_mat_device(A, W, jlo, jhi, af=af)
        assert b == pos and torch.equal(part, whole[b:e]), (jlo, jhi)
        pos = e
        del part
    assert pos == whole.numel()


# ------------------------------------------------------------------------------------------------------------
# C5: rank 4 (x) rank 4, dim 40, fp32 -> rank 8 (314,457,495 components), then the vector contraction
# ------------------------------------------------------------------------------------------------------------
def _outer_components(af, bf, ra, rb, dim, K):
    """Packed formula of multiply.outer (oracle/packed_oracle.py::outer) for the sorted multi-indices K [n, ra+rb];
    returns (values, sums of |terms|)."""
    n = ra + rb
    subsets = list(itertools.combinations(range(n), ra))
    acc = np.zeros(K.shape[0])
    mag = np.zeros(K.shape[0])
    for S in subsets:
        Sc = [q for q in range(n) if q not in S]
        t = af[po.flat_rank_array(K[:, list(S)], dim)] * bf[po.flat_rank_array(K[:, Sc], dim)]
        acc += t
        mag += np.abs(t)
    return acc / len(subsets), mag / len(subsets)


def test_config5_outer_windows_and_identity():
    ra = rb = 4
    dim = 40
    A = device_tensor(ra, dim, 20261018 + 5, torch.float32)
    B = device_tensor(rb, dim, 20261018 + 55, torch.float32)
    x = np.random.default_rng(5).uniform(0.5, 1.5, dim).astype(np.float32) / np.float32(np.sqrt(dim))
    C = st.multiply.outer(A, B)
    assert C.rank == 8 and C.dim == dim and C.dtype == np.float32 and C.indep_size == 314457495
    af, bf = flat_host(A), flat_host(B)  # fp32 values, up-cast
    table = C.class_table
    rng = np.random.default_rng(55)
    c_host = C.packed.cpu().numpy()
    nchecked = 0
    for cls in table.classes:
        ci = table.index(cls)
        pos = sample_positions(table.sizes[ci], 200, rng)
        if len(pos) == 0:
            continue
        K = sorted_indices(cls, dim, pos)
        want, mag = _outer_components(af, bf, ra, rb, dim, K)
        got = c_host[table.offsets[ci] + pos].astype(np.float64)
        assert np.all(np.abs(got - want) <= 1e-5 * mag), (cls, np.max(np.abs(got - want) / mag))
        nchecked += len(pos)
    assert nchecked > 3000
    del c_host
    # (A (x)_s B) . x^8 == (A . x^4)(B . x^4): the stored rank-8 tensor through the vector kernel, and the fused kernel
    rhs = float(st.contract_all_indices_with_vector(A.astype(np.float64), x.astype(np.float64))) * \
        float(st.contract_all_indices_with_vector(B.astype(np.float64), x.astype(np.float64)))
    unfused = float(st.contract_all_indices_with_vector(C, x))
    fused = float(ops.outer_then_contract_vec(A, B, x))
    assert abs(unfused - rhs) <= 1e-5 * abs(rhs), (unfused, rhs)
    assert abs(fused - rhs) <= 1e-5 * abs(rhs), (fused, rhs)


# ------------------------------------------------------------------------------------------------------------
# C3: rank 3 dim 1000 fp32 tensordot over one index
# ------------------------------------------------------------------------------------------------------------
def _tensordot_components(af, bf, ra, rb, dim, K):
    """Packed formula of tensordot with ONE contracted index (oracle/packed_oracle.py::tensordot, k = 1) for the sorted
    multi-indices K [n, ra+rb-2]; returns (values, sums of |terms|)."""
    n = ra + rb - 2
    na = ra - 1
    subsets = list(itertools.combinations(range(n), na))
    acc = np.zeros(K.shape[0])
    mag = np.zeros(K.shape[0])
    a_all = np.arange(dim, dtype=np.int64)
    for S in subsets:
        Sc = [q for q in range(n) if q not in S]
        for row in range(K.shape[0]):
            ia = np.sort(np.concatenate([np.broadcast_to(K[row, list(S)], (dim, na)), a_all[:, None]], axis=1), axis=1)
            ib = np.sort(np.concatenate([np.broadcast_to(K[row, Sc], (dim, n - na)), a_all[:, None]], axis=1), axis=1)
            t = af[po.flat_rank_array(ia, dim)] * bf[po.flat_rank_array(ib, dim)]
            acc[row] += t.sum()
            mag[row] += np.abs(t).sum()
    return acc / len(subsets), mag / len(subsets)


@pytest.fixture(scope="module")
def c3_operands():
    return device_tensor(3, 1000, 20261018 + 3, torch.float32), device_tensor(3, 1000, 20261018 + 33, torch.float32)


def test_config3_ladder_rung_windows(c3_operands):
    """The (ra, rb, k) = (3, 2, 1) rung of SURVEY.md 8d's ladder at dim 1000 (output rank 3, 167 M components)."""
    A, _ = c3_operands
    dim = 1000
    B = device_tensor(2, dim, 20261018 + 32, torch.float32)
    C = st.tensordot(A, B, axes=1)
    assert C.rank == 3 and C.dim == dim and C.dtype == np.float32
    af, bf = flat_host(A), flat_host(B)
    table = C.class_table
    c_host = C.packed.cpu().numpy()
    rng = np.random.default_rng(33)
    for cls in table.classes:
        ci = table.index(cls)
        pos = sample_positions(table.sizes[ci], 60, rng)
        K = sorted_indices(cls, dim, pos)
        want, mag = _tensordot_components(af, bf, 3, 2, dim, K)
        got = c_host[table.offsets[ci] + pos].astype(np.float64)
        assert np.all(np.abs(got - want) <= 1e-5 * mag), (cls, np.max(np.abs(got - want) / mag))


def test_config3_full_operands_output_ranges_windows(c3_operands):
    """BASELINE config 3 itself -- rank 3 . rank 3 over one index at dim 1000 (output rank 4, 41,917,125,250 components =
    167.7 GB, which only exists sharded): the tiled tcgen05 kernel computes RANGES of the packed output from the full-size
    operands, exactly what one GPU of an 8-GPU run does for its slice; random windows of every range against the packed
    formula in fp64 (1e-5 of the sum of |terms|).  Ranges: the start of the buffer (the four small classes and the first
    components of class (1,1,1,1)), a slice from the middle, and the very end."""
    from symtensor_b200._cabi import c_i64, lib
    A, B = c3_operands
    dim = 1000
    assert lib.st_tensordot_is_tiled(3, 3, 1, c_i64(dim), 4) == 1
    table = comb.class_table(4, dim)
    total = table.total
    af_d, bf_d = ops._flat_buffer(A, torch.float32), ops._flat_buffer(B, torch.float32)
    af, bf = af_d.cpu().numpy().astype(np.float64), bf_d.cpu().numpy().astype(np.float64)
    rng = np.random.default_rng(333)
    n_small = table.offsets[4]                      # start of class (1,1,1,1)
    span = 200_000_000 // 32 * 32                   # 0.8 GB of output per range
    mid = (total // 2) // 32 * 32
    ws = None
    for begin, end in [(0, n_small + span), (mid, mid + span), (total - span, total)]:
        out = torch.empty(end - begin, dtype=torch.float32, device=DEV)
        ws = ops.tensordot_device(A, B, 1, out, begin, end, torch.float32, af=af_d, bf=bf_d, ws=ws)
        for ci, cls in enumerate(table.classes):
            lo, hi = max(begin, table.offsets[ci]), min(end, table.offsets[ci] + table.sizes[ci])
            if lo >= hi:
                continue
            pos = np.unique(np.concatenate([np.arange(lo, min(hi, lo + 4)), np.arange(max(lo, hi - 4), hi), rng.integers(lo, hi, 24)]))
            K = sorted_indices(cls, dim, pos - table.offsets[ci])
            want, mag = _tensordot_components(af, bf, 3, 3, dim, K)
            got = out[torch.as_tensor(pos - begin, device=DEV)].cpu().numpy().astype(np.float64)
            assert np.all(np.abs(got - want) <= 1e-5 * mag), (begin, cls, np.max(np.abs(got - want) / mag))
        # alignment padding of the packed layout stays zero
        for ci in range(table.ncls):
            lo, hi = max(begin, table.offsets[ci] + table.sizes[ci]), min(end, table.offsets[ci + 1])
            if lo < hi:
                assert float(out[lo - begin:hi - begin].abs().max()) == 0.0
        del out


def test_config3_shard_of_an_8_gpu_partition_in_one_call(c3_operands):
    """What one GPU of an 8-GPU run of BASELINE config 3 computes (sharding.tensordot22_shards: its part of EVERY class with
    repeated indices plus its part of class (1,1,1,1), five disjoint ranges in one st_tensordot_ranges_f32 call -- the tiles
    that hold a diagonal are spread over the GPUs).  Shards 0 and 5 restricted to 60 M components of their last range; windows
    of every range against the packed formula in fp64; the ranges of all shards tile the buffer."""
    A, B = c3_operands
    dim = 1000
    table = comb.class_table(4, dim)
    shards = sharding.tensordot22_shards(dim, 8)
    covered = sorted(r for sh in shards for r in sh)
    assert covered[0][0] == 0 and covered[-1][1] == table.total
    assert all(a[1] <= b[0] and b[0] - a[1] < 32 for a, b in zip(covered[:-1], covered[1:]))  # disjoint; gaps are alignment padding only
    af_d, bf_d = ops._flat_buffer(A, torch.float32), ops._flat_buffer(B, torch.float32)
    af, bf = af_d.cpu().numpy().astype(np.float64), bf_d.cpu().numpy().astype(np.float64)
    rng = np.random.default_rng(3333)
    ws = None
    for g in (0, 5):
        ranges = list(shards[g])
        assert 2 <= len(ranges) <= 5
        b4, e4 = ranges[-1]
        ranges[-1] = (b4, min(e4, b4 + 60_000_000 // 32 * 32))
        outs = [torch.full((e - b,), -7.0, dtype=torch.float32, device=DEV) for b, e in ranges]
        ws = ops.tensordot_device_ranges(A, B, 1, outs, ranges, af=af_d, bf=bf_d, ws=ws)
        for (begin, end), out in zip(ranges, outs):
            for ci, cls in enumerate(table.classes):
                lo, hi = max(begin, table.offsets[ci]), min(end, table.offsets[ci] + table.sizes[ci])
                if lo >= hi:
                    continue
                pos = np.unique(np.concatenate([np.arange(lo, min(hi, lo + 4)), np.arange(max(lo, hi - 4), hi), rng.integers(lo, hi, 16)]))
                K = sorted_indices(cls, dim, pos - table.offsets[ci])
                want, mag = _tensordot_components(af, bf, 3, 3, dim, K)
                got = out[torch.as_tensor(pos - begin, device=DEV)].cpu().numpy().astype(np.float64)
                assert np.all(np.abs(got - want) <= 1e-5 * mag), (g, begin, cls, np.max(np.abs(got - want) / mag))
        del outs

