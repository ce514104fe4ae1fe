"""Measurement of BASELINE configs[2..4] (configs 3, 4, 5 in SURVEY.md's numbering) for ``bench.py``: each returns a dict for the
JSON line -- time, packed components/s, the roofline that bounds the op -- at N = 1 (one GPU does the slice an 8-GPU run would
give it where the whole does not fit) and at N > 1 (every rank its slice, MAX of the device times over the ranks; no data-path
collective except the scalar all-reduce after the fused outer -> vector).  Inputs are seeded synthetic tensors generated on the
device (values U[0.5, 1.5)).  Host logic only; the timed work is the C-ABI library's kernels.  (``cpu_legs`` is the one function here that runs the
oracle: it is the CPU baseline leg of the bench, not part of any timed GPU path.)"""
from __future__ import annotations

import math

import numpy as np
import torch

from symtensor_b200 import combinatorics as comb
from symtensor_b200 import ops, sharding
from symtensor_b200._cabi import c_i64, lib
from symtensor_b200.permcls import CudaPermClsSymmetricTensor

SEED = 20261018


def device_tensor(rank, dim, seed, dtype, dev):
    t = comb.class_table(rank, dim)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    buf = (torch.rand(t.total, generator=g, dtype=torch.float64, device=dev) + 0.5).to(dtype)
    for i in range(t.ncls):
        buf[t.offsets[i] + t.sizes[i]:t.offsets[i + 1]] = 0
    return CudaPermClsSymmetricTensor.from_packed(rank, dim, buf)


def _timed(fn, reps, warm, dist=None):
    for _ in range(warm):
        fn()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.st_launch_count()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms, int(lib.st_launch_count() - l0) // reps


def config3(dev, world, rank, dist, peaks, share_of=8):
    """rank 3 . rank 3 over one index, dim 1000, fp32 -> rank 4 (41,917,125,250 components, 167.7 GB): the output only exists
    sharded.  N = 1: this GPU computes share `share_of // 2` of a `share_of`-way partition (tile-aligned, balanced by tile count);
    N > 1: every rank computes its share of an N-way partition."""
    dim = 1000
    A, B = device_tensor(3, dim, SEED + 3, torch.float32, dev), device_tensor(3, dim, SEED + 33, torch.float32, dev)
    af, bf = ops._flat_buffer(A, torch.float32), ops._flat_buffer(B, torch.float32)
    parts = world if world > 1 else share_of
    # a GPU's shard: its part of every class with repeated indices + its part of class (1,1,1,1) (sharding.tensordot22_shards),
    # computed by ONE call over five output ranges
    me = rank if world > 1 else parts // 2
    ranges = sharding.tensordot22_shards(dim, parts)[me]
    outs = [torch.empty(e - b, dtype=torch.float32, device=dev) for b, e in ranges]
    state = {"ws": None}

    def step():
        state["ws"] = ops.tensordot_device_ranges(A, B, 1, outs, ranges, af=af, bf=bf, ws=state["ws"], check_flag=False)
    ms, launches = _timed(step, 2, 1, dist if world > 1 else None)
    flag = int(state["ws"][:1].view(torch.int32)[0].item())
    total = sum(comb.class_table(4, dim).sizes)
    tab = comb.class_table(4, dim)
    n_done = total if world > 1 else sum(min(e, o + s) - max(b, o) for b, e in ranges for o, s in zip(tab.offsets, tab.sizes) if min(e, o + s) > max(b, o))
    flops = 12.0 * dim * n_done           # 2 d C(4, 2) per packed output component, fp32-equivalent
    pipe = 3.0 * flops / world            # 3xTF32: three tensor-core products per fp32 product, per GPU
    peak = peaks["bf16_tflops"] / 2.0     # kind::tf32 runs at half the bf16 rate
    ach = pipe / (ms * 1e-3) / 1e12
    res = {"workload": f"tensordot rank 3 . rank 3 over one index, dim {dim}, fp32 -> rank 4 (BASELINE configs[2]); "
                       + (f"every rank its shard of a {world}-way tile-balanced partition of the 167.7 GB output (five ranges per rank)" if world > 1 else
                          f"one GPU: shard {me} of a {parts}-way tile-balanced partition of the 167.7 GB output (what one of {parts} GPUs computes)"),
           "packed_components": int(n_done), "ms": ms, "value": n_done / (ms * 1e-3), "unit": "packed components/s",
           "tflops_fp32_equivalent": flops / (ms * 1e-3) / 1e12, "kernel_launches": launches, "error_flag": flag,
           "includes": "expansion of both operands into hi/lo pair matrices, zeroing of the output range, the tiled tcgen05 kernel",
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "peak_source": "half of MEASURED_PEAKS.json bf16_tflops (kind::tf32 issues at half the bf16 rate)" if peaks["measured"]
                        else "half of the fallback bf16 figure of B200_PROFILING.md",
                        "kernel": "sym22_umma_kernel (tcgen05.mma kind::tf32, TMA-staged 4-D operand boxes, 3xTF32: per-GPU tensor-pipe flops = 3 x 12 d per component)"}}
    del outs, state
    return res


def config4(dev, world, rank, dist, peaks):
    """rank 6 dim 64 fp64 contract_all_indices_with_matrix (119,877,472 components, 2.534e12 algorithmic flops)."""
    r, dim = 6, 64
    A = device_tensor(r, dim, SEED + 4, torch.float64, dev)
    g = torch.Generator(device="cpu")
    g.manual_seed(SEED + 44)
    W = ((torch.rand(dim, dim, generator=g, dtype=torch.float64) + 0.5) / dim).numpy()
    n = comb.indep_size(r, dim)
    flops = sum(2 * dim * math.comb(dim + k - 1, k) * dim * math.comb(dim + r - k - 2, r - k - 1) for k in range(r))
    af = ops._flat_buffer(A, torch.float64)
    if world == 1:
        ms, launches = _timed(lambda: ops._contract_all_indices_with_matrix(A, W), 2, 1)
        what = "whole tensor on one GPU"
    else:
        cuts = sharding.mat_mode_bounds(r, dim, world)
        jlo, jhi = cuts[rank], cuts[rank + 1]
        ms, launches = _timed(lambda: ops.contract_mat_device(A, W, jlo, jhi, af=af), 2, 1, dist)
        what = f"partition by the first output mode over {world} GPUs (cuts {cuts}); the result stays sharded"
    peak = 40.0
    ach = flops / world / (ms * 1e-3) / 1e12 if world > 1 else flops / (ms * 1e-3) / 1e12
    # what the kernels multiply: only the 8-column blocks that hold a column j >= max(J) (sharding.mat_mode_work, in tiles of
    # 64 rows x 8 columns x dim) -- a third of SURVEY.md 8d's count, which charges all dim columns for every J
    mult = sum(sharding.mat_mode_work(r, dim)) * 64 * 8 * dim * 2
    return {"workload": f"contract_all_indices_with_matrix rank {r} dim {dim} fp64, W {dim} x {dim} (BASELINE configs[3]); " + what,
            "packed_components": n, "ms": ms, "value": n / (ms * 1e-3), "unit": "packed components/s", "kernel_launches": launches,
            "roofline": {"bound": "fp64 tensor pipe", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                         "peak_source": "nominal B200 FP64 (DMMA) rate, 40 TFLOP/s (tcgen05 has no f64 kind; MEASURED_PEAKS.json has no fp64 entry)",
                         "kernel": "mat_pipe_kernel / mat_last_dmma_kernel (persistent cp.async producer / DMMA consumer pipeline, mma.sync m8n8k4 f64 "
                                   "mode chain); algorithmic flops 2.534e12 (SURVEY.md 8d: all dim columns for every J)"
                                   + ("; per-GPU average of the partition" if world > 1 else ""),
                         "flops_multiplied": mult, "achieved_multiplied_tflops": mult / world / (ms * 1e-3) / 1e12}}


def config5(dev, world, rank, dist, peaks):
    """multiply.outer of two rank-4 dim-40 fp32 tensors -> rank 8 (314,457,495 components), then the vector contraction:
    unfused (store the rank-8 tensor, stream it) and fused (never stored); output range sharded for N > 1."""
    ra = rb = 4
    dim = 40
    A, B = device_tensor(ra, dim, SEED + 5, torch.float32, dev), device_tensor(rb, dim, SEED + 55, torch.float32, dev)
    af, bf = ops._flat_buffer(A, torch.float32), ops._flat_buffer(B, torch.float32)
    g = torch.Generator(device="cpu")
    g.manual_seed(SEED + 5)
    x = ((torch.rand(dim, generator=g, dtype=torch.float32) + 0.5) / dim ** 0.5).to(dev)
    table = comb.class_table(ra + rb, dim)
    n = sum(table.sizes)
    cuts = sharding.shard_bounds(table.total, world)
    d = dist if world > 1 else None
    balance = None
    if world > 1:
        # the cost per coordinate varies along the packed range (the small classes at the start: short rows, more class
        # changes): every rank times its own slice and the cuts move until the slices take the same time (sharding.rebalance)
        for _ in range(2):
            b, e = cuts[rank], cuts[rank + 1]
            tmp = torch.empty(max(e - b, 1), dtype=torch.float32, device=dev)
            for _w in range(2):
                ops.outer_device(A, B, tmp, b, e, torch.float32, af=af, bf=bf)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.outer_device(A, B, tmp, b, e, torch.float32, af=af, bf=bf)
            e1.record()
            torch.cuda.synchronize()
            tl = torch.zeros(world, dtype=torch.float64, device=dev)
            tl[rank] = e0.elapsed_time(e1)
            dist.all_reduce(tl)
            times = [float(v) for v in tl.cpu()]
            cuts = sharding.rebalance(cuts, times)
            balance = {"slice_ms_before_last_round": [round(v, 3) for v in times],
                       "slice_fractions": [round((cuts[r + 1] - cuts[r]) / table.total, 4) for r in range(world)]}
            del tmp
    b, e = cuts[rank], cuts[rank + 1]
    out = torch.empty(e - b, dtype=torch.float32, device=dev)
    ms_outer, l1 = _timed(lambda: ops.outer_device(A, B, out, b, e, torch.float32, af=af, bf=bf), 2, 1, d)
    res = torch.zeros(1, dtype=torch.float32, device=dev)
    ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)

    class _D:
        layout, rank, dim = 0, ra + rb, 40
    dsc = _D()
    dsc._buf = out

    def vec_step():
        ops.contract_vec_device(dsc, x, res, ws, b, e, packed=out)
        if world > 1:
            dist.all_reduce(res)
    ms_vec, l2 = _timed(vec_step, 3, 1, d)
    unfused = float(res[0])
    fws = torch.empty(int(lib.st_outer_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)
    fres = torch.zeros(1, dtype=torch.float32, device=dev)
    from symtensor_b200.ops import _fn, _stream_ptr, check

    def fused_step():
        check(_fn("st_outer_vec", torch.float32)(ra, rb, c_i64(dim), af.data_ptr(), bf.data_ptr(), x.data_ptr(), fres.data_ptr(), fws.data_ptr(),
                                                 c_i64(b), c_i64(e), _stream_ptr(dev)))
        if world > 1:
            dist.all_reduce(fres)
    ms_fused, l3 = _timed(fused_step, 2, 1, d)
    fused = float(fres[0])
    xa = float(ops._contract_all_indices_with_vector(A, x)) * float(ops._contract_all_indices_with_vector(B, x))
    # The kernel writes 4 B per component and does 70 products of two GATHERED operand entries for it (the operands, 494 KB
    # each, live in L1 / L2): its ceiling is the rate at which an SM's load pipe takes warp-wide loads -- one instruction (32
    # lanes) per clock per SM when the lanes fall into one or two cache lines, which consecutive components mostly do.
    sms, clk = torch.cuda.get_device_properties(dev).multi_processor_count, 1.965e9
    peak = sms * clk * 32 / 1e9
    ach = 140.0 * n / world / (ms_outer * 1e-3) / 1e9
    return {"workload": f"multiply.outer rank 4 (x) rank 4 dim {dim} fp32 -> rank 8, then contract_all_indices_with_vector (BASELINE configs[4]); "
                        + ("one GPU" if world == 1 else f"output range sharded over {world} GPUs, scalar all-reduce after the vector contraction"),
            "packed_components": n, "ms_outer": ms_outer, "ms_vector": ms_vec, "ms_fused_outer_vector": ms_fused,
            "ms": ms_outer + ms_vec, "value": n / ((ms_outer + ms_vec) * 1e-3), "unit": "packed components/s",
            "value_fused": n / (ms_fused * 1e-3), "gathers_per_s": 140.0 * n / (ms_outer * 1e-3),
            "hbm_written_gbs_per_gpu": n * 4 / world / (ms_outer * 1e-3) / 1e9,
            "result_unfused": unfused, "result_fused": fused, "identity_(A.x^4)(B.x^4)": xa,
            "rel_err_unfused": abs(unfused - xa) / abs(xa), "rel_err_fused": abs(fused - xa) / abs(xa),
            "kernel_launches": l1 + l2 + l3, "balance": balance,
            "roofline": {"bound": "l1 load pipe (gathered operand loads)", "achieved": ach, "peak": peak, "unit": "G lane-loads/s", "frac": ach / peak,
                         "traffic": None,
                         "peak_source": f"{sms} SMs x 1.965 GHz x 32 lanes: one warp-wide load per clock per SM (B300_MICROARCH.md); HBM is not the bound -- "
                                        "4 B are written per component (hbm_written_gbs_per_gpu)",
                         "kernel": "outer_rows_kernel<float,4,4> (warp-uniform row walk, lanes decode their component, 70 products = 140 gathered "
                                   "operand loads per output component; ncu: 37 warp instructions per component, issue slots 56 % busy, "
                                   "profiles/r2_outer_rows_ncu_full.txt)"}}


def cpu_legs():
    """The packed NumPy oracle (a restatement of the reference formulas, oracle/packed_oracle.py) on reduced dimensions: the
    reference itself cannot run configs 3-5 at any size close to these (dense arrays of 4 TB / 550 GB / 26 TB)."""
    import time

    from oracle import index_oracle as io
    from oracle import packed_oracle as po
    rng = np.random.default_rng(SEED)

    def rp(rank, dim):
        return {c: rng.uniform(0.5, 1.5, io.permclass_size(c, dim)) for c in io.perm_classes(rank)}
    out = {}
    d3 = 40
    A, B = rp(3, d3), rp(3, d3)
    t0 = time.perf_counter()
    po.tensordot(A, 3, B, 3, d3, 1)
    dt = time.perf_counter() - t0
    out["config3"] = {"value": io.indep_size(4, d3) / dt, "unit": "packed components/s", "cores": 1, "kind": "port",
                      "sample": f"packed NumPy oracle, tensordot rank 3 . rank 3 over one index at dim {d3} fp64 ({io.indep_size(4, d3)} comps, {dt:.1f} s)"}
    r4, d4 = 6, 12
    A = rp(r4, d4)
    W = rng.uniform(0.5, 1.5, (d4, d4)) / d4
    t0 = time.perf_counter()
    po.contract_all_indices_with_matrix(A, r4, d4, W)
    dt = time.perf_counter() - t0
    out["config4"] = {"value": io.indep_size(r4, d4) / dt, "unit": "packed components/s", "cores": 1, "kind": "port",
                      "sample": f"packed NumPy oracle, matrix contraction rank {r4} dim {d4} fp64 ({io.indep_size(r4, d4)} comps, {dt:.1f} s)"}
    d5 = 12
    A, B = rp(4, d5), rp(4, d5)
    t0 = time.perf_counter()
    C = po.outer(A, 4, B, 4, d5)
    po.contract_all_indices_with_vector(C, 8, d5, rng.uniform(0.5, 1.5, d5))
    dt = time.perf_counter() - t0
    out["config5"] = {"value": io.indep_size(8, d5) / dt, "unit": "packed components/s", "cores": 1, "kind": "port",
                      "sample": f"packed NumPy oracle, outer rank 4 (x) rank 4 + vector contraction at dim {d5} fp64 ({io.indep_size(8, d5)} comps, {dt:.1f} s)"}
    return out
