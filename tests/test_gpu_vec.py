"""contract_all_indices_with_vector on the GPU, through the reference-facing API (symtensor_b200.symalg) and the
C-ABI, against: the unmodified reference's outputs (tests/golden), the packed / C oracles on seeded inputs, and
size-independent properties at BASELINE.json's full sizes.

Tolerances (north-star): 1e-12 relative for fp64, 1e-5 relative for fp32 -- relative to the sum of |terms| for
signed inputs (condition number), relative to the value itself for the all-positive headline
distribution."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import index_oracle as io
from oracle import packed_oracle as po

import symtensor_b200 as st
from symtensor_b200 import combinatorics as comb
from symtensor_b200 import ops
from symtensor_b200._cabi import c_i64, check, lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL64, RTOL32 = 1e-12, 1e-5


def abs_terms(data, rank, dim, x):
    return po.contract_all_indices_with_vector({k: np.abs(v) for k, v in data.items()}, rank, dim, np.abs(x))


@pytest.fixture(autouse=True)
def _reset_variant():
    """variant 0 = production (small classes per component, the rest through the tail tables), 1 = generic
    per-component kernel, 2 = every class through the tail tables (so small test tensors exercise them)."""
    yield
    lib.st_set_vec_variant(0)


def make(rank, dim, seed, dist="pos", dtype=np.float64):
    rng = np.random.default_rng(seed)
    data = {}
    for c in io.perm_classes(rank):
        n = io.permclass_size(c, dim)
        data[c] = (rng.uniform(0.5, 1.5, n) if dist == "pos" else rng.standard_normal(n)).astype(dtype)
    x = (rng.uniform(0.5, 1.5, dim) / np.sqrt(dim) if dist == "pos" else rng.standard_normal(dim)).astype(dtype)
    return data, x


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_reference_goldens(goldens, variant):
    check(lib.st_set_vec_variant(variant))
    for c in goldens.cases("vec"):
        data, x, ref = goldens.packed(c["tag"] + ".A"), goldens.ops[c["tag"] + ".x"], float(goldens.ops[c["tag"] + ".out"])
        A = st.PermClsTorchSymmetricTensor(rank=c["rank"], dim=c["dim"], data=dict(data), device=DEV)
        res = st.contract_all_indices_with_vector(A, x)
        assert isinstance(res, st.PermClsTorchSymmetricTensor) and res.rank == 0 and res.dim == 1
        full = {k: np.broadcast_to(v, (io.permclass_size(k, c["dim"]),)) for k, v in data.items()}
        assert abs(float(res) - ref) <= RTOL64 * abs_terms(full, c["rank"], c["dim"], x), c


def test_config1_matches_the_reference_run(goldens):
    """BASELINE config 1: rank 4 dim 50 fp64, value computed by the reference itself (6.6 s there)."""
    c = goldens.config1
    rng = np.random.default_rng(c["seed"])
    data = {cls: rng.uniform(0.5, 1.5, io.permclass_size(cls, 50)) for cls in io.perm_classes(4)}
    x = rng.uniform(0.5, 1.5, 50) / np.sqrt(50)
    A = st.PermClsTorchSymmetricTensor(rank=4, dim=50, data=data, device=DEV)
    for variant in (0, 1, 2):
        check(lib.st_set_vec_variant(variant))
        got = float(st.contract_all_indices_with_vector(A, x))
        assert abs(got - c["result"]) <= RTOL64 * abs(c["result"])


@pytest.mark.parametrize("rank,dim", [(1, 33), (2, 40), (3, 21), (4, 16), (5, 11), (6, 10), (7, 9), (8, 10), (2, 300), (3, 64)])
@pytest.mark.parametrize("dist", ["pos", "normal"])
def test_against_packed_oracle_fp64(rank, dim, dist):
    data, x = make(rank, dim, 100 * rank + dim, dist)
    ref = po.contract_all_indices_with_vector(data, rank, dim, x)
    scale = abs_terms(data, rank, dim, x)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device=DEV)
    for variant in (0, 1, 2):
        check(lib.st_set_vec_variant(variant))
        assert abs(float(st.contract_all_indices_with_vector(A, x)) - ref) <= RTOL64 * scale


@pytest.mark.parametrize("rank,dim", [(2, 50), (3, 30), (4, 20), (6, 9), (8, 8)])
def test_fp32_against_fp64_oracle(rank, dim):
    """fp32 parity is taken against the fp64 oracle on the up-cast inputs (the reference's own fp32 path raises
    for most ops)."""
    data, x = make(rank, dim, 7 * rank + dim, "pos", np.float32)
    ref = po.contract_all_indices_with_vector({k: v.astype(np.float64) for k, v in data.items()}, rank, dim, x.astype(np.float64))
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device=DEV)
    assert A.dtype == np.float32
    for variant in (0, 1, 2):
        check(lib.st_set_vec_variant(variant))
        res = st.contract_all_indices_with_vector(A, x)
        assert res.dtype == np.float32
        assert abs(float(res) - ref) <= RTOL32 * abs(ref)
    # dtype promotion like NumPy: fp32 tensor with fp64 x gives fp64
    res = st.contract_all_indices_with_vector(A, x.astype(np.float64))
    assert res.dtype == np.float64 and abs(float(res) - ref) <= 1e-6 * abs(ref)


def test_scalar_classes_and_missing_classes(goldens):
    c = [q for q in goldens.cases("vec") if q.get("dist") == "scalar-classes"][0]
    data, x, ref = goldens.packed(c["tag"] + ".A"), goldens.ops[c["tag"] + ".x"], float(goldens.ops[c["tag"] + ".out"])
    assert any(np.ndim(v) == 0 for v in data.values())
    A = st.PermClsTorchSymmetricTensor(rank=4, dim=6, data=dict(data), device=DEV)
    full = {k: np.broadcast_to(v, (io.permclass_size(k, 6),)) for k, v in data.items()}
    assert abs(float(st.contract_all_indices_with_vector(A, x)) - ref) <= RTOL64 * abs_terms(full, 4, 6, x)
    # classes absent from the dict are zeros
    B = st.PermClsTorchSymmetricTensor(rank=3, dim=5, data={"iij": np.ones(20)}, device=DEV)
    x5 = np.arange(1.0, 6.0)
    ref = po.contract_all_indices_with_vector({(3,): np.zeros(5), (2, 1): np.ones(20), (1, 1, 1): np.zeros(10)}, 3, 5, x5)
    assert abs(float(st.contract_all_indices_with_vector(B, x5)) - ref) <= RTOL64 * abs(ref)


def test_reference_error_and_early_exit_behaviour():
    A = st.PermClsTorchSymmetricTensor(rank=3, dim=4, data=1.0, device=DEV)
    with pytest.raises(ValueError, match="must match"):
        st.contract_all_indices_with_vector(A, np.ones(5))
    r = st.contract_all_indices_with_vector(A, np.zeros(4))
    assert r == 0 and isinstance(r, int)  # symtensor/symalg.py:519-520
    # A = all ones  =>  (sum x)^3
    x = np.array([0.5, -1.0, 2.0, 3.0])
    assert abs(float(st.contract_all_indices_with_vector(A, x)) - x.sum() ** 3) <= 1e-12 * np.abs(x).sum() ** 3
    # rank-3 dim-3 dense truth as in symtensor/testing/api.py:657-672
    rng = np.random.default_rng(3)
    dense = rng.standard_normal((3, 3, 3))
    dense = sum(dense.transpose(p) for p in [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]) / 6
    T = st.PermClsTorchSymmetricTensor(data=dense, device=DEV)
    x3 = rng.standard_normal(3)
    assert np.isclose(float(st.contract_all_indices_with_vector(T, x3)), np.einsum("abc,a,b,c->", dense, x3, x3, x3))


def test_flat_layout(goldens):
    for c in goldens.cases("flatvec"):
        v, x, ref = goldens.ops[c["tag"] + ".A"], goldens.ops[c["tag"] + ".x"], float(goldens.ops[c["tag"] + ".out"])
        F = st.FlatSymmetricTensor(c["rank"], c["dim"], v, device=DEV)
        res = st.contract_all_indices_with_vector(F, x)
        assert isinstance(res, st.FlatSymmetricTensor) and res.rank == 0
        assert abs(float(res) - ref) <= RTOL64 * abs(ref)
    for rank, dim in [(3, 25), (5, 9), (8, 7)]:
        rng = np.random.default_rng(rank)
        v = rng.standard_normal(comb.indep_size(rank, dim))
        x = rng.standard_normal(dim)
        ref = po.contract_vec_flat(v, rank, dim, x)
        scale = po.contract_vec_flat(np.abs(v), rank, dim, np.abs(x))
        F = st.FlatSymmetricTensor(rank, dim, v, device=DEV)
        assert abs(float(st.contract_all_indices_with_vector(F, x)) - ref) <= RTOL64 * scale
        # same tensor in the permcls layout gives the same number
        P = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=po.flat_to_permcls(v, rank, dim), device=DEV)
        assert abs(float(st.contract_all_indices_with_vector(P, x)) - ref) <= RTOL64 * scale


def _raw_partial(A, xd, begin, end):
    out = torch.zeros(1, dtype=A.torch_dtype, device=DEV)
    ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=DEV)
    ops.contract_vec_device(A, xd, out, ws, begin, end, packed=A.packed[begin:end])
    return float(out[0])


@pytest.mark.parametrize("rank,dim", [(4, 60), (6, 14), (8, 11)])
def test_range_partials_sum_to_the_whole(rank, dim):
    """The [begin, end) interface used for multi-GPU sharding and host streaming: partial sums over any
    32-aligned split add up to the full contraction; each shard only sees its own slice of the buffer."""
    data, x = make(rank, dim, rank + dim)
    ref = po.contract_all_indices_with_vector(data, rank, dim, x)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device=DEV)
    xd = torch.as_tensor(x, device=DEV)
    total = A.packed.numel()
    for variant in (0, 1, 2):
        check(lib.st_set_vec_variant(variant))
        for nshards in (2, 3, 8):
            cuts = [min(total, (total * i // nshards + 31) // 32 * 32) for i in range(nshards)] + [total]
            s = sum(_raw_partial(A, xd, b, e) for b, e in zip(cuts[:-1], cuts[1:]))
            assert abs(s - ref) <= RTOL64 * abs(ref)
    with pytest.raises(ValueError):
        _raw_partial(A, xd, 16, total)  # begin must be 32-aligned


def test_host_buffers_stream_through_the_gpu():
    """End-to-end entry with HOST buffers (st_contract_vec_host_*): chunked H2D copies overlap the kernel."""
    rank, dim = 4, 110  # 6.3M comps = 50 MB -> several 64 MiB chunks? no: one; use fp64 dim 160 for >1 chunk below
    data, x = make(rank, dim, 5)
    ref = po.contract_all_indices_with_vector(data, rank, dim, x)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device="host")
    res = st.contract_all_indices_with_vector(A, x)
    assert abs(float(res) - ref) <= RTOL64 * abs(ref)
    A32 = A.astype(np.float32)
    assert abs(float(st.contract_all_indices_with_vector(A32, x.astype(np.float32))) - ref) <= RTOL32 * abs(ref)


def test_config2_full_size_against_c_oracle():
    """BASELINE config 2 at full size: rank 4 dim 200 fp64, 68,685,050 packed components, against the C oracle
    (long-double accumulation, literal σindex_iter loops); plus host streaming (9 chunks), linearity and
    determinism."""
    rank, dim = 4, 200
    data, x = make(rank, dim, 20261018 + 2)
    ref = co.contract_all_indices_with_vector(data, rank, dim, x)
    A = st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=data, device=DEV)
    assert A.size == 68685050
    got = float(st.contract_all_indices_with_vector(A, x))
    assert abs(got - ref) <= RTOL64 * abs(ref)
    assert float(st.contract_all_indices_with_vector(A, x)) == got  # fixed reduction order => bit reproducible
    check(lib.st_set_vec_variant(1))
    assert abs(float(st.contract_all_indices_with_vector(A, x)) - ref) <= RTOL64 * abs(ref)
    check(lib.st_set_vec_variant(0))
    # homogeneity: A . (2x)^4 = 16 A . x^4 exactly (powers of two)
    assert float(st.contract_all_indices_with_vector(A, 2.0 * x)) == 16.0 * got
    # host-resident tensor streamed in chunks
    Ah = A.to("host")
    assert abs(float(st.contract_all_indices_with_vector(Ah, x)) - ref) <= RTOL64 * abs(ref)
    # fp32 at full size against the fp64 oracle
    A32 = A.astype(np.float32)
    r32 = float(st.contract_all_indices_with_vector(A32, x.astype(np.float32)))
    ref32 = co.contract_all_indices_with_vector({k: v.astype(np.float32).astype(np.float64) for k, v in data.items()}, rank, dim,
                                                x.astype(np.float32).astype(np.float64))
    assert abs(r32 - ref32) <= RTOL32 * abs(ref32)


@pytest.mark.parametrize("rank,dim,dtype", [(8, 40, np.float32), (6, 64, np.float64)])
def test_large_rank_configs_via_identities(rank, dim, dtype):
    """Rank 8 dim 40 (config 5's contraction, 314M comps) and rank 6 dim 64 (config 4's tensor, 120M comps):
    beyond any dense method; checked with A = all-ones => (sum x)^r and variant-1 vs variant-0 agreement on a
    structured tensor."""
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    t = comb.class_table(rank, dim)
    buf = torch.ones(t.total, dtype=tdt, device=DEV)
    for c, s, o in zip(t.classes, t.sizes, t.offsets):  # keep the alignment padding zero
        buf[o + s:t.offsets[t.index(c) + 1]] = 0
    A = st.PermClsTorchSymmetricTensor.from_packed(rank, dim, buf)
    rng = np.random.default_rng(rank)
    x = (rng.uniform(0.5, 1.5, dim) / dim).astype(dtype)
    expect = float(np.sum(x.astype(np.float64))) ** rank
    tol = RTOL32 if dtype == np.float32 else RTOL64
    got = float(st.contract_all_indices_with_vector(A, x))
    assert abs(got - expect) <= tol * expect
    # position-dependent values: both kernels must agree (they share no index code)
    buf.mul_(torch.linspace(0.5, 1.5, t.total, dtype=tdt, device=DEV))
    a = float(st.contract_all_indices_with_vector(A, x))
    check(lib.st_set_vec_variant(1))
    b = float(st.contract_all_indices_with_vector(A, x))
    assert abs(a - b) <= tol * abs(b)


@pytest.mark.parametrize("rank,dim,dtype", [(4, 200, np.float64), (4, 50, np.float64), (6, 30, np.float64), (3, 400, np.float32), (5, 9, np.float64)])
def test_overlapped_batch_equals_single_launches(rank, dim, dtype):
    """ST_VEC_OVERLAP (programmatic dependent launch: launch i + 1 starts while launch i drains its tail, alternating
    workspace halves and claim-counter sets): a batch of contractions of one resident tensor with n vectors gives, bit for
    bit, the results of n isolated launches -- the result does not depend on the tile deal -- and matches the oracle."""
    t = comb.class_table(rank, dim)
    g = torch.Generator(device=DEV)
    g.manual_seed(rank * 1000 + dim)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    buf = (torch.rand(t.total, generator=g, dtype=torch.float64, device=DEV) + 0.5).to(tdt)
    for i in range(t.ncls):
        buf[t.offsets[i] + t.sizes[i]:t.offsets[i + 1]] = 0
    A = st.PermClsTorchSymmetricTensor.from_packed(rank, dim, buf)
    rng = np.random.default_rng(dim)
    n = 9
    X = (rng.uniform(0.5, 1.5, (n, dim)) / np.sqrt(dim)).astype(dtype)
    batch = ops.contract_all_indices_with_vectors(A, X).cpu().numpy()
    single = np.array([float(st.contract_all_indices_with_vector(A, X[i])) for i in range(n)], dtype=dtype)
    assert np.array_equal(batch, single)
    batch2 = ops.contract_all_indices_with_vectors(A, X).cpu().numpy()  # and is reproducible run to run
    assert np.array_equal(batch, batch2)
    if t.total <= 2_000_000:
        host = {c: A._data[c].cpu().numpy().astype(np.float64) for c in t.classes}
        for i in (0, n - 1):
            ref = po.contract_all_indices_with_vector(host, rank, dim, X[i].astype(np.float64))
            assert abs(float(batch[i]) - ref) <= (RTOL64 if dtype == np.float64 else RTOL32) * abs(ref)
    with pytest.raises(ValueError):
        ops.contract_all_indices_with_vectors(A, X[:, :-1])
