"""Host logic of the tiled tensordot kernel (st_sym22.cu): the tile list of an output range [begin, end) -- the multi-GPU
partition of BASELINE config 3 -- against a brute-force enumeration with the ORACLE's (class, position) map: every tile that
holds a component of the range must be in the list (a missing tile is a missing part of the result)."""
import itertools

import numpy as np

from oracle import index_oracle as io

from symtensor_b200 import combinatorics as comb
from symtensor_b200 import sharding
from symtensor_b200._cabi import c_i64, lib


def tiles_needed(dim, begin, end):
    tab = comb.class_table(4, dim)
    need = set()
    for t in itertools.combinations_with_replacement(range(dim), 4):
        cls, pos = io.rank_of_index(t, dim)
        if begin <= tab.offsets[tab.index(cls)] + pos < end:
            need.add((t[0] // 16, t[1] // 16, t[2] // 16, t[3] // 8))
    return need


def tiles_listed(dim, begin, end):
    buf = np.zeros(1 << 16, dtype=np.uint64)
    n = lib.st_debug_sym22_tiles(c_i64(dim), c_i64(begin), c_i64(end), buf.ctypes.data, c_i64(buf.size))
    assert 0 <= n <= buf.size
    words = [int(w) for w in buf[:n]]
    assert len(set(words)) == n  # no tile twice: a component gets its three adds from ONE tile
    return {(w & 0xffff, (w >> 16) & 0xffff, (w >> 32) & 0xffff, (w >> 48) & 0xffff) for w in words}


def test_tile_lists_cover_every_output_range():
    for dim in (7, 17, 33, 40):
        total = comb.class_table(4, dim).total
        for world in (1, 2, 3, 8):
            cuts = sharding.shard_bounds(total, world)
            union = set()
            for b, e in zip(cuts[:-1], cuts[1:]):
                need, got = tiles_needed(dim, b, e), tiles_listed(dim, b, e)
                assert need <= got, (dim, b, e, sorted(need - got)[:5])
                union |= got
            assert union == tiles_listed(dim, 0, total) == tiles_needed(dim, 0, total)
    assert tiles_listed(40, 64, 64) == set()


def test_tile_lists_of_ranges_inside_the_small_classes_are_pruned_but_complete():
    """Ranges that cut the classes with repeated indices (the two-range shards of sharding.tensordot22_shards): the list still
    holds every tile that the range needs, and a narrow range no longer drags in every diagonal tile."""
    rng = np.random.default_rng(5)
    for dim in (17, 40):
        tab = comb.class_table(4, dim)
        off4 = tab.offsets[tab.ncls - 1]
        cutsets = [sorted({0, off4} | {int(v) // 32 * 32 for v in rng.integers(0, off4, size=5)}) for _ in range(2)]
        for cuts in cutsets:
            for b, e in zip(cuts[:-1], cuts[1:]):
                need, got = tiles_needed(dim, b, e), tiles_listed(dim, b, e)
                assert need <= got, (dim, b, e, sorted(need - got)[:5])
        if dim >= 40:
            allt = tiles_listed(dim, 0, off4)
            b = tab.offsets[3] + (tab.sizes[3] // 2) // 32 * 32  # a narrow window in the middle of class (2,1,1)
            assert len(tiles_listed(dim, b, b + 64)) < len(allt)
    import ctypes
    for dim, world in ((40, 3), (33, 4)):
        shards = sharding.tensordot22_shards(dim, world)
        total = comb.class_table(4, dim).total
        covered = sorted(r for sh in shards for r in sh if r[1] > r[0])
        assert covered[0][0] == 0 and covered[-1][1] == total
        assert all(a[1] <= b[0] and b[0] - a[1] < 32 for a, b in zip(covered[:-1], covered[1:]))  # disjoint; gaps: alignment padding
        for sh in shards:  # the tile list of a shard (all its ranges in one call) holds what each of its ranges needs, once
            n = len(sh)
            bs, es = (ctypes.c_int64 * n)(*[r[0] for r in sh]), (ctypes.c_int64 * n)(*[r[1] for r in sh])
            buf = np.zeros(1 << 16, dtype=np.uint64)
            cnt = lib.st_debug_sym22_tiles_ranges(c_i64(dim), n, bs, es, buf.ctypes.data, c_i64(buf.size))
            words = [int(w) for w in buf[:cnt]]
            assert len(set(words)) == cnt
            got = {(w & 0xffff, (w >> 16) & 0xffff, (w >> 32) & 0xffff, (w >> 48) & 0xffff) for w in words}
            for b, e in sh:
                assert tiles_needed(dim, b, e) <= got


def test_config3_shards_are_balanced_by_the_tile_cost_model():
    """sharding.tensordot22_shards at BASELINE config 3 (dim 1000, 8 GPUs): at most five ranges per GPU, and the modelled cost of the
    tiles each GPU runs (tiles + 3.2 sum 1 / (l blocks of the tile's k block), fitted to measured shard times) within 8 % of
    the mean -- one contiguous range per GPU put 1.9x the mean on the first."""
    import ctypes
    dim, world = 1000, 8
    shards = sharding.tensordot22_shards(dim, world)
    costs = []
    for ranges in shards:
        assert 1 <= len(ranges) <= 5  # (a GPU whose part of class (1,1,1,1) is already above the mean gets no small-class interval)
        n = len(ranges)
        bs, es = (ctypes.c_int64 * n)(*[r[0] for r in ranges]), (ctypes.c_int64 * n)(*[r[1] for r in ranges])
        stats = (ctypes.c_double * 2)()
        assert lib.st_debug_sym22_tiles_stats(c_i64(dim), n, bs, es, stats) == 0
        costs.append(stats[0] + 3.2 * stats[1])
    assert max(costs) <= 1.08 * sum(costs) / world, costs
