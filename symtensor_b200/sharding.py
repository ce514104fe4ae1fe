"""Multi-GPU partitioning of the hot path (one process per GPU, ``torch.distributed``).

The path shards without any data-path collective:

* ``contract_all_indices_with_vector``: the packed coordinate range is cut into ``world`` contiguous 32-aligned
  slices balanced by bytes; every rank holds only its slice, ``x`` is replicated, and the ONLY collective is the
  sum all-reduce of one scalar (NCCL over NVLink on GPUs, gloo in the CPU tests);
* ``multiply.outer`` / ``tensordot`` / fused outer->vector: the OUTPUT packed range is cut the same way, the (small)
  operands are replicated, the tensor result stays sharded -- no collective (a scalar all-reduce if a vector
  contraction follows).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch

ALIGN = 32  # ST_CLASS_ALIGN: shard boundaries keep 16-byte vector loads aligned


def shard_bounds(total: int, world: int, align: int = ALIGN) -> List[int]:
    """``world + 1`` cut points of ``[0, total)``: contiguous slices, starts aligned, sizes balanced."""
    if world < 1:
        raise ValueError("world must be >= 1")
    cuts = [min(total, (total * i // world + align - 1) // align * align) for i in range(world)]
    return cuts + [total]


def rebalance(cuts: List[int], times: List[float], align: int = ALIGN) -> List[int]:
    """New cut points from the measured time of every slice: the cost per coordinate differs along the packed range
    (classes with earlier runs and the long first rows of the big class cost several times the table walk), so
    slices of equal bytes are not slices of equal time.  The cost density is taken as constant inside each old
    slice; the new cuts split the cumulative cost evenly.  Pure host arithmetic (unit-tested on the CPU)."""
    world = len(cuts) - 1
    if world != len(times) or world < 1:
        raise ValueError("need one time per slice")
    total_t = float(sum(times))
    if world == 1 or total_t <= 0:
        return list(cuts)
    new = [cuts[0]]
    acc = 0.0  # cost up to cuts[r]
    r = 0
    for k in range(1, world):
        target = total_t * k / world
        while r < world - 1 and acc + times[r] < target:
            acc += times[r]
            r += 1
        length = cuts[r + 1] - cuts[r]
        frac = (target - acc) / times[r] if times[r] > 0 else 0.0
        x = cuts[r] + int(length * min(max(frac, 0.0), 1.0))
        x = min(cuts[-1], max(new[-1], (x + align - 1) // align * align))
        new.append(x)
    return new + [cuts[-1]]


def my_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    cuts = shard_bounds(total, world)
    return cuts[rank], cuts[rank + 1]


def all_reduce_sum(value: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank partial results (in place); a no-op outside a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(value, op=dist.ReduceOp.SUM, group=group)
    return value


def contract_vec_sharded(rank_: int, dim: int, shard: torch.Tensor, x: torch.Tensor, begin: int, end: int, out: torch.Tensor,
                         ws: Optional[torch.Tensor] = None, group=None, partial_fn: Optional[Callable] = None, async_op: bool = False,
                         overlap: bool = False, partial_only: bool = False):
    """Vector contraction of a range-sharded permcls tensor: local streaming kernel over ``[begin, end)`` (``shard``
    starts at coordinate ``begin``), then the scalar all-reduce.  ``partial_fn(shard, x, begin, end) -> float`` replaces
    the CUDA launch in the CPU (gloo) tests of the host logic.  With ``async_op`` the all-reduce is only enqueued (on the
    collective's own stream, ordered after the kernel) and its work handle is returned: the caller may launch the next
    contraction -- into another ``out`` -- before waiting, so that the collective overlaps the next kernel.  ``overlap``
    (``ST_VEC_OVERLAP``, resident operands only): the kernel starts while the previous launch on the stream drains its tail;
    ``ws`` must then hold twice ``st_contract_vec_workspace_bytes()``.  ``partial_only``: only the local kernel -- the caller
    reduces several partial sums with one collective (``out`` may be a one-element view into a vector of partial sums)."""
    if partial_fn is not None:
        out.fill_(partial_fn(shard, x, begin, end))
    else:
        from . import ops

        class _Desc:
            layout = 0
        d = _Desc()
        d.rank, d.dim, d._buf = rank_, dim, shard
        if ws is None:
            from ._cabi import lib
            ws = torch.empty((2 if overlap else 1) * int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64,
                             device=shard.device)
        ops.contract_vec_device(d, x, out, ws, begin, end, packed=shard, overlap=overlap)
    if partial_only:
        return None
    if async_op:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group, async_op=True)
        return None
    return all_reduce_sum(out, group)


def tensordot22_bounds(dim: int, world: int, align: int = ALIGN) -> List[int]:
    """Cut points of the packed rank-4 output of a tensordot with two free indices per side (BASELINE config 3) for ``world``
    GPUs: contiguous ranges whose boundaries sit where the FIRST index enters a new block of 16 -- the tiles of the tcgen05
    kernel (st_sym22.cu) are 16 x 16 x 16 x 8 index blocks, and a range boundary inside a block makes both neighbours run the
    block's tiles -- balanced by the number of tiles each range has to run (``st_debug_sym22_tiles``: rank 0 also runs every
    tile that holds components with repeated indices, which live in the small classes at the start of the buffer).  No
    collective: the result stays sharded.  Pure host arithmetic."""
    import ctypes
    import math

    from . import combinatorics as comb
    from ._cabi import c_i64, lib

    table = comb.class_table(4, dim)
    total = table.total
    if world <= 1:
        return [0, total]
    off4 = table.offsets[table.ncls - 1]
    # first component whose smallest index is i0: (i0, i0+1, i0+2, i0+3) in class (1,1,1,1), lexicographic rank
    def first_coord(i0):
        r = math.comb(dim, 4) - 1 - (math.comb(dim - 1 - i0, 4) + math.comb(dim - 2 - i0, 3) + math.comb(dim - 3 - i0, 2) + math.comb(dim - 4 - i0, 1))
        return off4 + r
    cands = [0] + [min(total, first_coord(i0) // align * align) for i0 in range(16, dim - 3, 16)] + [total]
    cands = sorted(set(cands))

    def ntiles(b, e):
        return int(lib.st_debug_sym22_tiles(c_i64(dim), c_i64(b), c_i64(e), ctypes.c_void_p(0), c_i64(0))) if e > b else 0

    cuts = [0]
    lo_idx = 0
    for r in range(world - 1):
        remaining = ntiles(cuts[-1], total)
        target = remaining / (world - r)
        # smallest candidate whose range reaches the target (tile counts grow with the end point), then the closer neighbour
        lo, hi = lo_idx + 1, len(cands) - 1 - (world - 2 - r)
        hi = max(hi, lo)
        a, b = lo, hi
        while a < b:
            mid = (a + b) // 2
            if ntiles(cuts[-1], cands[mid]) >= target:
                b = mid
            else:
                a = mid + 1
        best = a
        if a > lo and abs(ntiles(cuts[-1], cands[a - 1]) - target) <= abs(ntiles(cuts[-1], cands[a]) - target):
            best = a - 1
        best = min(best, len(cands) - 1)
        cuts.append(cands[best])
        lo_idx = best
    return cuts + [total]


def tensordot22_shards(dim: int, world: int, align: int = ALIGN) -> List[List[tuple]]:
    """Output ranges per GPU for a tensordot with two free indices per side on ``world`` GPUs, balanced by the number of tiles
    each GPU runs: ``shards[g]`` is a list of disjoint ``(begin, end)`` ranges of the packed rank-4 output, computed by ONE call
    of ``ops.tensordot_device_ranges`` (``st_tensordot_ranges_f32``).

    ``tensordot22_bounds`` gives every GPU one contiguous range; the components with repeated indices live in the small
    classes at the start of the buffer, so GPU 0 runs every tile that holds a diagonal on top of its share of class
    (1,1,1,1) (dim 1000, 8 GPUs: 318,060 tiles against 159,000 - 211,000 for the others -- the step time is GPU 0's: 713 ms
    against ~400).  Here GPU g takes (i) the components of EVERY small class whose first class-order value a -- a repeated
    index -- lies in its interval ``[a_g, a_g+1)``: four ranges that need the same diagonal tiles (those whose overlapping
    index blocks meet the interval), the intervals cut where the tile count of the prefix reaches g / world; (ii) its part
    of class (1,1,1,1), cut where the first index enters a new block of 16 so that small + large tile COSTS are even.
    No collective: the result stays sharded, five pieces per GPU.  Pure host arithmetic."""
    import ctypes
    import math

    from . import combinatorics as comb
    from ._cabi import c_i64, lib

    table = comb.class_table(4, dim)
    total = table.total
    if table.classes != ((4,), (3, 1), (2, 2), (2, 1, 1), (1, 1, 1, 1)):
        raise RuntimeError("unexpected class order of rank 4")
    off = table.offsets
    off4 = off[4]
    if world <= 1:
        return [[(0, total)]]

    def ntiles(ranges):
        """Cost of the tiles the ranges need, in tiles: a tile whose k block r leaves few l blocks shares its three column boxes
        with few neighbours and costs more -- measured on the eight shards of dim 1000 (tools/shard_times_c3.py):
        ms = 0.001995 tiles + 0.006384 sum 1 / (l blocks of r), residuals below 1 % (the last shard: 2.41 us per tile against 2.15)."""
        ranges = [r for r in ranges if r[1] > r[0]]
        if not ranges:
            return 0.0
        n = len(ranges)
        b = (ctypes.c_int64 * n)(*[r[0] for r in ranges])
        e = (ctypes.c_int64 * n)(*[r[1] for r in ranges])
        stats = (ctypes.c_double * 2)()
        if lib.st_debug_sym22_tiles_stats(c_i64(dim), n, b, e, stats) != 0:
            raise RuntimeError("st_debug_sym22_tiles_stats failed")
        return stats[0] + 3.2 * stats[1]

    # first position of the value a in each small class (a = dim: the end of the class)
    first = [lambda a: a,
             lambda a: a * (dim - 1),
             lambda a: a * (dim - 1) - a * (a - 1) // 2,
             lambda a: a * math.comb(dim - 1, 2)]

    def small_ranges(a0, a1):
        out = []
        for c in range(4):
            b = off[c] + (first[c](a0) // align * align if a0 > 0 else 0)
            e = off[c] + (first[c](a1) // align * align if a1 < dim else table.sizes[c])
            e = min(e, off[c] + table.sizes[c])
            out.append((b, max(b, e)))
        return out

    # class (1,1,1,1) first: cut where the first index enters a new block of 16 (a boundary inside a block would make both
    # neighbours run the block's tiles), as even as that coarse grid allows
    def first_coord(i0):
        r = math.comb(dim, 4) - 1 - (math.comb(dim - 1 - i0, 4) + math.comb(dim - 2 - i0, 3) + math.comb(dim - 3 - i0, 2) + math.comb(dim - 4 - i0, 1))
        return off4 + r
    cands = sorted(set([off4] + [min(total, max(off4, first_coord(i0) // align * align)) for i0 in range(16, dim - 3, 16)] + [total]))
    big = [off4]
    idx = 0
    big_all = ntiles([(off4, total)])
    for g in range(1, world):
        target = big_all * g / world  # cost of the prefix [off4, cut): errors of the coarse grid do not pile up on the last GPU
        lo, hi = idx, max(idx, len(cands) - 1 - (world - 1 - g))
        a, b = lo, hi
        while a < b:
            mid = (a + b) // 2
            if ntiles([(off4, cands[mid])]) >= target:
                b = mid
            else:
                a = mid + 1
        best = a
        if a > lo and abs(ntiles([(off4, cands[a - 1])]) - target) <= abs(ntiles([(off4, cands[a])]) - target):
            best = a - 1
        best = min(best, len(cands) - 1)
        big.append(cands[best])
        idx = best
    big.append(total)
    n_big = [ntiles([(big[g], big[g + 1])]) for g in range(world)]

    # then the intervals of a: GPU g gets the diagonal tiles that bring it to the common total (the grid of a is fine: dim values)
    d_all = ntiles(small_ranges(0, dim))
    per_gpu = (d_all + sum(n_big)) / world
    want = [max(0.0, per_gpu - n) for n in n_big]
    scale = d_all / max(sum(want), 1.0)
    want = [w * scale for w in want]
    acuts = [0]
    for g in range(world - 1):
        lo, hi = acuts[-1], dim
        while lo < hi:
            mid = (lo + hi) // 2
            if ntiles(small_ranges(acuts[-1], mid)) >= want[g]:
                hi = mid
            else:
                lo = mid + 1
        acuts.append(min(max(lo, acuts[-1]), dim))
    acuts.append(dim)
    small = [small_ranges(acuts[g], acuts[g + 1]) for g in range(world)]
    return [[r for r in small[g] if r[1] > r[0]] + ([(big[g], big[g + 1])] if big[g + 1] > big[g] else []) for g in range(world)]


def mat_mode_work(rank: int, dim: int) -> List[float]:
    """Tensor-pipe work of the matrix contraction's mode chain per value j1 of the FIRST output mode, in tiles of 64 rows x 8
    columns x dim: step k multiplies, for every row J = (j1 <= ... <= jk) and every tile of 64 I, only the 8-column blocks
    that hold a column j >= jk (st_mat.cu), so the work of a slice is far from proportional to SURVEY.md 8d's flop count (all
    dim columns for every J), which round 1 balanced.  Step 0 (J empty) is charged per column."""
    import math

    def flat(k):
        return math.comb(dim + k - 1, k) if k > 0 else 1

    def count(j1, jl, k):  # sorted k-tuples over range(dim) that start with j1 and end with jl
        if k == 1:
            return 1 if jl == j1 else 0
        return math.comb(jl - j1 + k - 2, k - 2) if jl >= j1 else 0
    nblk_all = (dim + 7) // 8
    work = []
    for j1 in range(dim):
        c = math.ceil(flat(rank - 1) / 64) / 8.0 if rank >= 1 else 0.0
        for k in range(1, rank):
            m = rank - k - 1
            tiles = math.ceil(flat(m) / 64) if m > 0 else 1.0 / 64
            c += tiles * sum(count(j1, jl, k) * (nblk_all - jl // 8) for jl in range(j1, dim))
        work.append(c)
    return work


def mat_mode_bounds(rank: int, dim: int, world: int) -> List[int]:
    """Cut points ``j_0 = 0 < j_1 < ... < j_world = dim`` of the FIRST output mode for the matrix contraction on ``world`` GPUs
    (``ops.contract_mat_device``): GPU g computes the output components whose smallest index lies in ``[j_g, j_{g+1})``.  Balanced
    by the tensor-pipe work the slice's mode chain actually does (``mat_mode_work``).  Pure host arithmetic."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if rank < 1 or dim < 1:
        return [0] + [dim] * world
    work = mat_mode_work(rank, dim)
    total = sum(work)
    cuts = [0]
    acc = 0.0
    j = 0
    for g in range(1, world):
        target = total * g / world
        while j < dim and acc + work[j] < target:
            acc += work[j]
            j += 1
        if j < dim and abs(acc + work[j] - target) < abs(acc - target):
            acc += work[j]
            j += 1
        cuts.append(max(j, cuts[-1]))
    return cuts + [dim]
