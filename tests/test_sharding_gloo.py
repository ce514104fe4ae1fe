"""Host-side logic of the multi-GPU path on the CPU: world_size-2 ``gloo`` process group, range partition of the
packed coordinates, scalar all-reduce.  The per-rank partial sums come from the oracle here (test infrastructure);
on GPUs the same function launches the CUDA kernel (tests/test_gpu_vec.py covers the range interface itself)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import index_oracle as io
from oracle import packed_oracle as po

from symtensor_b200 import combinatorics as comb
from symtensor_b200 import sharding


def test_shard_bounds_are_aligned_contiguous_and_balanced():
    for total in (0, 32, 64, 1000 * 32, 68685952, 12345 * 32):
        for world in (1, 2, 3, 4, 8):
            cuts = sharding.shard_bounds(total, world)
            assert cuts[0] == 0 and cuts[-1] == total and len(cuts) == world + 1
            assert all(a <= b for a, b in zip(cuts[:-1], cuts[1:]))
            assert all(c % 32 == 0 for c in cuts[:-1])
            sizes = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
            if total >= 64 * world:
                assert max(sizes) - min(sizes) <= 64
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 0)


def _weights(rank, dim, x):
    """gamma_c * prod x^m per packed coordinate of the padded permcls buffer (oracle side)."""
    t = comb.class_table(rank, dim)
    w = np.zeros(t.total)
    for c, s, o in zip(t.classes, t.sizes, t.offsets):
        if s == 0:
            continue
        vals = io.class_values(c, dim)
        ww = np.full(s, float(io.permclass_multiplicity(c)))
        for k, m in enumerate(c):
            ww *= x[vals[:, k]] ** m
        w[o:o + s] = ww
    return w


def _worker(rank, world, port, rank_t, dim, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        data = {c: rng.uniform(0.5, 1.5, io.permclass_size(c, dim)) for c in io.perm_classes(rank_t)}
        x = rng.uniform(0.5, 1.5, dim)
        t = comb.class_table(rank_t, dim)
        buf = np.zeros(t.total)
        for c, s, o in zip(t.classes, t.sizes, t.offsets):
            buf[o:o + s] = data[c]
        w = _weights(rank_t, dim, x)
        begin, end = sharding.my_range(t.total, rank, world)
        shard = torch.from_numpy(buf[begin:end].copy())  # every rank holds only its slice

        def partial(sh, xx, b, e):
            return float(np.dot(sh.numpy(), w[b:e]))
        out = torch.zeros(1, dtype=torch.float64)
        sharding.contract_vec_sharded(rank_t, dim, shard, torch.from_numpy(x), begin, end, out, group=None, partial_fn=partial)
        ref = po.contract_all_indices_with_vector(data, rank_t, dim, x)
        q.put((rank, float(out[0]), ref, end - begin))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rank_t,dim", [(4, 12), (6, 7)])
def test_range_sharded_vector_contraction_world2_gloo(rank_t, dim):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rank_t, dim, 123, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = comb.class_table(rank_t, dim).total
    assert sum(r[3] for r in res) == total
    for _, got, ref, _ in res:  # every rank holds the all-reduced value
        assert abs(got - ref) <= 1e-12 * abs(ref)


def test_rebalance_equalises_cost():
    """sharding.rebalance: slices of equal measured time from slices of equal bytes (pure host arithmetic)."""
    from symtensor_b200 import sharding
    total = 1 << 20
    cuts = sharding.shard_bounds(total, 4)
    # cost density 3 on the first quarter of the range, 1 elsewhere
    dens = lambda a, b: 3.0 * max(0, min(b, total // 4) - a) + 1.0 * max(0, b - max(a, total // 4))  # noqa: E731
    for _ in range(4):
        times = [dens(cuts[r], cuts[r + 1]) for r in range(4)]
        cuts = sharding.rebalance(cuts, times)
        assert cuts[0] == 0 and cuts[-1] == total and all(c % 32 == 0 for c in cuts[:-1]) and cuts == sorted(cuts)
    times = [dens(cuts[r], cuts[r + 1]) for r in range(4)]
    assert max(times) / (sum(times) / 4) < 1.02
    assert sharding.rebalance([0, 100], [1.0]) == [0, 100]


def test_matrix_mode_bounds_balance_the_chain_work():
    """sharding.mat_mode_bounds (host arithmetic): monotone cuts covering [0, dim]; at BASELINE config 4 (rank 6 dim 64, 8 GPUs)
    no slice carries more than 1.15x the mean tensor-pipe work of the chain (sharding.mat_mode_work: only the 8-column blocks
    with a column j >= max(J) are multiplied; the first mode is an integer: perfect balance is not available)."""
    from symtensor_b200 import sharding
    for rank, dim, world in [(6, 64, 8), (4, 12, 3), (3, 20, 1), (2, 5, 8), (6, 64, 2), (1, 7, 3)]:
        cuts = sharding.mat_mode_bounds(rank, dim, world)
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == dim
        assert all(a <= b for a, b in zip(cuts[:-1], cuts[1:]))
    rank, dim, world = 6, 64, 8
    cuts = sharding.mat_mode_bounds(rank, dim, world)
    work = sharding.mat_mode_work(rank, dim)
    parts = [sum(work[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert max(parts) <= 1.15 * sum(parts) / world
    # the work of a first mode falls quickly with j1 (fewer rows J start there AND fewer columns lie above them)
    assert work[0] > 4 * work[dim // 2] > 0


def test_matrix_mode_work_matches_a_brute_force_count():
    """sharding.mat_mode_work against a direct enumeration of the items the step kernel runs: for every step k >= 1, every sorted
    k-tuple J that starts with j1, every tile of 64 I, the 8-column blocks that hold a column j >= max(J); step 0 per column."""
    import itertools
    import math

    from symtensor_b200 import sharding
    for rank, dim in [(3, 9), (4, 12), (5, 7), (2, 20), (4, 17)]:
        work = sharding.mat_mode_work(rank, dim)
        nblk = (dim + 7) // 8
        for j1 in (0, 1, dim // 2, dim - 1):
            want = math.ceil(math.comb(dim + rank - 2, rank - 1) / 64) / 8.0
            for k in range(1, rank):
                m = rank - k - 1
                tiles = math.ceil(math.comb(dim + m - 1, m) / 64) if m > 0 else 1.0 / 64
                for J in itertools.combinations_with_replacement(range(j1, dim), k - 1):
                    jl = J[-1] if J else j1
                    want += tiles * (nblk - jl // 8)
            assert abs(work[j1] - want) <= 1e-9 * max(1.0, want), (rank, dim, j1, work[j1], want)


def test_tensordot22_bounds_are_aligned_and_cover():
    from symtensor_b200 import combinatorics as comb
    from symtensor_b200 import sharding
    for dim, world in [(70, 2), (200, 4), (120, 8)]:
        total = comb.class_table(4, dim).total
        cuts = sharding.tensordot22_bounds(dim, world)
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == total
        assert all(a <= b for a, b in zip(cuts[:-1], cuts[1:])) and all(c % 32 == 0 for c in cuts[:-1])
