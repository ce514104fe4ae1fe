#!/usr/bin/env python
"""bench.py -- headline benchmark of the symmetrized-contraction hot path (BASELINE.json configs[1]):
``contract_all_indices_with_vector`` on a rank-4 dim-200 float64 permutation-class tensor (68,685,050 packed
components, 549.5 MB), reported as packed components/s plus the achieved fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full vector contraction of the resident packed tensor (ONE kernel launch: streaming pass with the
deterministic finalize fused in; for N > 1 followed by the scalar NCCL all-reduce).  N > 1 (launched by torchrun, one rank per GPU): WEAK scaling -- the
packed coordinate range of a rank-4 tensor whose dimension grows with N (200, 238, 283, 337: ~6.9e7 components
per GPU) is split into N contiguous 32-aligned slices, one per GPU; x is replicated; the only collective is the
all-reduce of the partial sums.  The strong-scaling figure (the dim-200 tensor itself cut N ways) is reported in
the extra key "strong".

JSON keys follow the driver contract; extra keys: roofline, cpu_baseline, cpu_packed_oracle, strong.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RANK, DIM = 4, 200
WEAK_DIMS = {1: 200, 2: 238, 4: 283, 8: 337}  # C(d+3, 4) ~ N * 68.7e6
SEED = 20261018 + 2
METRIC = "packed components/s (contract_all_indices_with_vector, permcls rank 4 fp64)"
CPU_SAMPLE_DIM = 112  # the reference's dense algorithm needs d**4 doubles: 12.8 GB at dim 200


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_peak_bf16():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["bf16_tflops"]), True
        except Exception:
            pass
    return 1590.0, False


def ncu_traffic(n_comps):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        if int(d.get("packed_components", -1)) == int(n_comps):
            return float(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def synth(rank, dim, seed):
    """Seeded synthetic tensor in the headline distribution: packed values ~ U[0.5, 1.5) generated
    per class in class order, x ~ U[0.5, 1.5) / sqrt(dim)."""
    from oracle import index_oracle as io
    rng = np.random.default_rng(seed)
    data = {c: rng.uniform(0.5, 1.5, io.permclass_size(c, dim)) for c in io.perm_classes(rank)}
    x = rng.uniform(0.5, 1.5, dim) / np.sqrt(dim)
    return data, x


# ----------------------------------------------------------------------------------------------------------
REF_CONFIG = (4, 50)  # BASELINE configs[0]: the reference's own CPU-runnable case (292,825 packed components, ~9 s per call)


def load_reference():
    """The UNMODIFIED reference package, staged into baseline/_ref by ``__graft_entry__.build()`` (it is pure Python; its four
    uninstallable dependencies are replaced by the stand-ins of oracle/ref_shim).  None when the copy is absent."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "symtensor")):
        return None
    import warnings
    warnings.filterwarnings("ignore")
    for p in (ref, os.path.join(ROOT, "oracle", "ref_shim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import symtensor  # noqa: F401
    from symtensor import symalg
    from symtensor.permcls_symtensor import PermClsSymmetricTensor
    return symalg, PermClsSymmetricTensor


def real_reference_arm(steps, warmup):
    """``symtensor.symalg.contract_all_indices_with_vector`` of the unmodified reference (NumPy backend) on BASELINE
    configs[0] (rank 4 dim 50 fp64): the largest case of this op the reference finishes in seconds -- configs[1] would take
    ~37 min and a 12.8 GB dense array per call.  Returns (comps/s, s per call, comps, value) or None."""
    loaded = load_reference()
    if loaded is None:
        return None
    symalg, PermCls = loaded
    r, d = REF_CONFIG
    data, x = synth(r, d, SEED - 1)
    A = PermCls(rank=r, dim=d, data={k: v.copy() for k, v in data.items()})
    n = sum(v.size for v in data.values())
    for _ in range(warmup):
        symalg.contract_all_indices_with_vector(A, x)
    times, val = [], None
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        val = symalg.contract_all_indices_with_vector(A, x)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return n / dt, dt, n, float(np.asarray(val._data[()]))


def cpu_reference_arm(steps, warmup):
    """A PORT of the reference's CPU algorithm for this path (densify -> np.tensordot -> r!-symmetrize -> repack, r
    times; symtensor/symalg.py:505-527), restated with vectorised NumPy in oracle/dense_oracle.py (~50x faster than the
    reference's Python loops).  Bounded sample: same op at dim 112."""
    from oracle import dense_oracle as do
    from oracle import index_oracle as io
    d = CPU_SAMPLE_DIM
    data, x = synth(RANK, d, SEED)
    n = io.indep_size(RANK, d)
    for _ in range(min(warmup, 1)):
        do.contract_all_indices_with_vector(data, RANK, d, x)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        do.contract_all_indices_with_vector(data, RANK, d, x)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return n / dt, dt, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(args.steps, 3)
    warm = min(args.warmup, 1)
    cores = os.cpu_count()
    real = real_reference_arm(steps, warm)
    if real is not None:
        value, dt, n, _ = real
        kind = "reference"
        used = 1
        sample = (f"the UNMODIFIED reference (symtensor.symalg.contract_all_indices_with_vector, PermClsSymmetricTensor, NumPy backend; "
                  f"one Python thread, BLAS only inside np.tensordot) on BASELINE configs[0]: rank 4 dim {REF_CONFIG[1]} fp64, {n} packed comps "
                  f"({n / 68685050:.2%} of configs[1], which would need ~37 min and a 12.8 GB dense array per call); {dt:.1f} s per call")
        port_v, port_dt, port_n = cpu_reference_arm(1, 0)
        port = {"value": port_v, "unit": "packed components/s", "cores": cores, "kind": "port",
                "sample": f"vectorised NumPy port of the same algorithm at dim {CPU_SAMPLE_DIM} ({port_n} packed comps, {port_dt:.1f} s)"}
    else:
        value, dt, n = cpu_reference_arm(steps, warm)
        kind, used, port = "port", cores, None
        sample = (f"baseline/_ref absent: NumPy port of the reference algorithm (dense d^4 array + np.tensordot + r! symmetrize + repack, x4) "
                  f"at dim {CPU_SAMPLE_DIM} ({n} packed comps, {n / 68685050:.1%} of the workload)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "packed components/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "rank 4 dim 200 float64 permcls contract_all_indices_with_vector (BASELINE configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "packed components/s", "cores": used, "kind": kind, "sample": sample,
                         "host_cores": cores},
        "e2e": {"value": value, "unit": "packed components/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if port is not None:
        line["cpu_port"] = port
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import symtensor_b200 as st
    from symtensor_b200 import combinatorics as comb
    from symtensor_b200 import ops
    from symtensor_b200._cabi import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL printf "NCCL version ..." to stdout: keep stdout for the ONE JSON line of the contract
        # (any other level the caller set -- e.g. INFO to see NVLS -- is left alone and goes where NCCL_DEBUG_FILE says)
        # (WARN prints the version line as well -- NCCL's showVersion() -- so the variable is dropped, not lowered)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)

    dim = WEAK_DIMS.get(world, DIM)
    table = comb.class_table(RANK, dim)
    n_comps = sum(table.sizes)
    total = table.total
    # contiguous 32-aligned slices of the packed coordinate range, balanced by bytes
    from symtensor_b200 import sharding
    cuts = sharding.shard_bounds(total, world)

    def make_shard(begin, end):
        # synthetic shard, generated on the device from a per-rank seed (values U[0.5, 1.5)); alignment padding zeroed
        g = torch.Generator(device=dev)
        g.manual_seed(SEED + rank)
        sh = torch.rand(end - begin, generator=g, dtype=torch.float64, device=dev) + 0.5
        for c in range(table.ncls):
            lo, hi = table.offsets[c] + table.sizes[c], table.offsets[c + 1]
            lo, hi = max(lo, begin), min(hi, end)
            if lo < hi:
                sh[lo - begin:hi - begin] = 0
        return sh

    begin, end = cuts[rank], cuts[rank + 1]
    shard = make_shard(begin, end)
    xg = torch.Generator(device="cpu")
    xg.manual_seed(SEED)
    x_host = (torch.rand(dim, generator=xg, dtype=torch.float64) + 0.5) / dim ** 0.5
    x = x_host.to(dev)
    A = st.PermClsTorchSymmetricTensor.from_packed(RANK, dim, shard) if world == 1 else None

    class _Desc:  # what ops.contract_vec_device needs to describe a sharded tensor
        layout, rank, dim = 0, RANK, None
    desc = _Desc()
    desc.dim = dim
    desc._buf = shard
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    # ST_VEC_OVERLAP launches alternate between the two halves of a double-size workspace
    ws = torch.empty(2 * int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)

    def balance(dim_, table_, cuts_, make, rounds=4):
        """Cost-balanced slices: the cost per coordinate varies along the packed range (classes with earlier runs, the long first
        rows of the big class), so every rank times its own kernel -- overlapped launches, as in the timed loop -- and the cut
        points are moved until the slices take the same time (sharding.rebalance).  Returns (cuts, shard, info)."""
        d_ = _Desc()
        d_.dim = dim_
        info = None
        sh = make(cuts_[rank], cuts_[rank + 1])
        for _ in range(rounds):
            d_._buf = sh
            b_, e_ = cuts_[rank], cuts_[rank + 1]
            for _ in range(3):
                ops.contract_vec_device(d_, x[:dim_], out, ws, b_, e_, packed=sh, overlap=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.contract_vec_device(d_, x[:dim_], out, ws, b_, e_, packed=sh, overlap=True)
            e1.record()
            torch.cuda.synchronize()
            tl = torch.zeros(world, dtype=torch.float64, device=dev)
            tl[rank] = e0.elapsed_time(e1) / 10
            dist.all_reduce(tl)
            times = [float(v) for v in tl.cpu()]
            cuts_ = sharding.rebalance(cuts_, times)
            del sh
            sh = make(cuts_[rank], cuts_[rank + 1])
            info = {"kernel_ms_per_rank_before_last_round": [round(v, 4) for v in times],
                    "slice_fractions": [round((cuts_[r + 1] - cuts_[r]) / table_.total, 4) for r in range(world)]}
        return cuts_, sh, info

    rebalanced = None
    if world > 1:
        del shard
        cuts, shard, rebalanced = balance(dim, table, cuts, make_shard)
        begin, end = cuts[rank], cuts[rank + 1]
        desc._buf = shard
    class Pipeline:
        """The timed loop over one sharded tensor.  Every step is one ST_VEC_OVERLAP kernel launch over this rank's slice (the
        operands are resident: the ramp-up of step i + 1 overlaps the tail of step i).  N > 1: the steps of a BLOCK write their
        partial sums into consecutive slots of one vector, and the block ends with ONE all-reduce of that vector (NCCL, enqueued
        asynchronously: it overlaps the kernels of the next block, which writes the other of two vectors) -- the path's only
        collective, "the final partial-sum all-reduce", paid once per block instead of once per step: an all-reduce kernel
        between two contractions takes SMs away from the next contraction's one-CTA-per-SM grid and costs more than it moves.
        The per-step host work (a ctypes call, ~10 us) would still hide kernels of a few tens of us on 8 GPUs, so a block is
        captured ONCE in a CUDA graph -- kernel launches with their programmatic edges, the all-reduce on the collective's
        stream -- and the timed region replays it.  Every step is still exactly one kernel launch."""

        def __init__(self, rank_, dim_, shard_, x_, begin_, end_, block):
            self.a = (rank_, dim_, shard_, x_, begin_, end_)
            self.block = max(1, block)
            self.outs = [torch.zeros(self.block, dtype=torch.float64, device=dev) for _ in range(2)]
            self.ws = ws
            self.pending = [None, None]
            self.count = 0
            self.graph, self.error = None, None

        def step(self):
            i, buf = self.count % self.block, (self.count // self.block) & 1
            self.count += 1
            r_, d_, sh_, x_, b_, e_ = self.a
            if i == 0 and self.pending[buf] is not None:
                self.pending[buf].wait()  # the vector is about to be rewritten: its all-reduce must be done
                self.pending[buf] = None
            sharding.contract_vec_sharded(r_, d_, sh_, x_, b_, e_, self.outs[buf][i:i + 1], self.ws, partial_only=True, overlap=True)
            if world > 1 and i == self.block - 1:
                self.pending[buf] = dist.all_reduce(self.outs[buf], op=dist.ReduceOp.SUM, async_op=True)

        def drain(self):
            for i in range(2):
                if self.pending[i] is not None:
                    self.pending[i].wait()
                    self.pending[i] = None

        def result(self):
            last = self.count - 1
            return float(self.outs[(last // self.block) & 1][last % self.block])

        def build_graph(self):
            if args.no_graph:
                return
            try:
                self.drain()
                torch.cuda.synchronize()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    self.count = 0
                    for _ in range(2 * self.block):  # warm-up on the capture stream (per-stream counters of the library, NCCL)
                        self.step()
                    self.drain()
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                self.count = 0
                with torch.cuda.graph(g, stream=side):
                    for _ in range(2 * self.block):  # both result vectors: a replay ends in the state it started in
                        self.step()
                    self.drain()
                torch.cuda.synchronize()
                g.replay()
                torch.cuda.synchronize()
                self.graph = g
            except Exception as ex:  # fall back to eager launches
                self.error = f"{type(ex).__name__}: {str(ex)[:200]}"
                self.graph = None
                self.pending = [None, None]
                try:
                    torch.cuda.synchronize()
                except Exception:
                    pass
            if world > 1:
                ok = torch.tensor([1 if self.graph is not None else 0], dtype=torch.int32, device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok[0]) == 0:  # all ranks or none
                    self.graph = None

        def timed(self, nsteps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = lib.st_launch_count()
            e0.record()
            if self.graph is not None and nsteps % (2 * self.block) == 0:
                for _ in range(nsteps // (2 * self.block)):
                    self.graph.replay()
                launched = nsteps
            else:
                for _ in range(nsteps):
                    self.step()
                self.drain()
                launched = lib.st_launch_count() - l0
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            return ms, launched

        def how(self):
            if self.graph is not None:
                return ("a CUDA graph of %d steps (kernel launches with programmatic edges%s), replayed" %
                        (2 * self.block, "; one NCCL all-reduce of %d partial sums per %d steps" % (self.block, self.block) if world > 1 else ""))
            return "one host launch per step" + (f" (graph capture failed: {self.error})" if self.error else "") + \
                ("; one NCCL all-reduce of %d partial sums per %d steps" % (self.block, self.block) if world > 1 else "")

    def graph_block(nsteps):
        for b in (10, 5, 4, 3, 2, 1):  # steps per all-reduce; a graph holds two blocks
            if nsteps % (2 * b) == 0:
                return b
        return 1

    pipe = Pipeline(RANK, dim, shard, x, begin, end, graph_block(args.steps))
    step, drain = pipe.step, pipe.drain

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    pipe.build_graph()
    with ClockSampler(local) as clocks:
        ms, launches = pipe.timed(args.steps)
    ms_per_step = ms / args.steps
    value = n_comps / (ms_per_step * 1e-3)
    # the same launches WITHOUT the overlap (each launch starts after the previous one has completely finished): the
    # latency of one isolated contraction, reported next to the pipelined figure
    isolated = None
    if world == 1:
        for _ in range(3):
            sharding.contract_vec_sharded(RANK, dim, shard, x, begin, end, out, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sharding.contract_vec_sharded(RANK, dim, shard, x, begin, end, out, ws)
        e1.record()
        torch.cuda.synchronize()
        iso_ms = e0.elapsed_time(e1) / args.steps
        isolated = {"ms_per_step": iso_ms, "value": n_comps / (iso_ms * 1e-3), "unit": "packed components/s",
                    "hbm_gbs": n_comps * 8 / (iso_ms * 1e-3) / 1e9,
                    "note": "back-to-back launches without ST_VEC_OVERLAP: launch i + 1 starts when launch i has completely finished"}
    result = pipe.result()
    serial = None
    if world > 1:
        # the same step without the overlap (kernel, then a blocking all-reduce): reported next to the headline
        def step_serial():
            sharding.contract_vec_sharded(RANK, dim, shard, x, begin, end, out, ws)
        for _ in range(3):
            step_serial()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_serial()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        serial = {"value": n_comps / (float(t[0]) / args.steps * 1e-3), "unit": "packed components/s",
                  "ms_per_step": float(t[0]) / args.steps, "note": "kernel, then blocking all-reduce (no overlap between steps)"}

    # ---- strong scaling: the dim-200 tensor itself cut `world` ways (reported, not the headline)
    strong = None
    if world > 1:
        t200 = comb.class_table(RANK, DIM)
        def make200(b_, e_):
            g2 = torch.Generator(device=dev)
            g2.manual_seed(SEED + 100 + rank)
            return torch.rand(e_ - b_, generator=g2, dtype=torch.float64, device=dev) + 0.5
        x2 = x[:DIM].contiguous()
        cuts2, sh2, bal2 = balance(DIM, t200, sharding.shard_bounds(t200.total, world), make200)
        b2, e2_ = cuts2[rank], cuts2[rank + 1]
        pipe2 = Pipeline(RANK, DIM, sh2, x2, b2, e2_, graph_block(args.steps))
        for _ in range(5):
            pipe2.step()
        pipe2.drain()
        pipe2.build_graph()
        ms2, _ = pipe2.timed(args.steps)
        t = torch.tensor([ms2], dtype=torch.float64, device=dev)
        strong = {"value": sum(t200.sizes) / (float(t[0]) / args.steps * 1e-3), "unit": "packed components/s",
                  "ms_per_step": float(t[0]) / args.steps, "workload": "rank 4 dim 200 cut %d ways" % world, "launch": pipe2.how(), "rebalance": bal2,
                  "note": "same pipeline as the headline (cost-balanced slices, overlapped launches, one all-reduce per block); "
                          "latency_ms_unpipelined is one isolated contraction: kernel, then a blocking all-reduce, launched from the host"}
        # un-pipelined latency of ONE contraction of the dim-200 tensor on N GPUs: kernel, then a blocking all-reduce, from the host
        for _ in range(3):
            sharding.contract_vec_sharded(RANK, DIM, sh2, x2, b2, e2_, out, ws)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sharding.contract_vec_sharded(RANK, DIM, sh2, x2, b2, e2_, out, ws)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong["latency_ms_unpipelined"] = float(t[0])
        del pipe2

    # ---- end to end through the reference-facing API with HOST buffers (rank 0 of N = 1 only)
    e2e = None
    cpu_baseline = None
    cpu_port = None
    cpu_packed = None
    if world == 1:
        Ah = A.to("host")  # pinned host copy of the packed tensor
        xh = x_host.numpy()
        n_e2e = max(3, min(10, args.steps))
        for _ in range(2):
            float(st.contract_all_indices_with_vector(Ah, xh))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            r_e2e = float(st.contract_all_indices_with_vector(Ah, xh))
        dt = (time.perf_counter() - t0) / n_e2e
        assert abs(r_e2e - result) <= 1e-12 * abs(result), (r_e2e, result)
        e2e = {"value": n_comps / dt, "unit": "packed components/s", "h2d_bytes_per_step": int(total * 8 + dim * 8),
               "d2h_bytes_per_step": 8, "ms_per_step": dt * 1e3,
               "api": "symtensor_b200.contract_all_indices_with_vector(host-resident PermClsTorchSymmetricTensor, x)"}
        del Ah
    else:
        # N > 1: every rank streams ITS slice from pinned host memory, runs the kernel, all-reduces, reads the scalar back
        sh_host = shard.cpu().pin_memory()
        sh_dev = torch.empty_like(shard)
        x_pin = x_host.pin_memory()
        x_dev = torch.empty_like(x)
        n_e2e = max(3, min(10, args.steps))

        def e2e_step():
            sh_dev.copy_(sh_host, non_blocking=True)
            x_dev.copy_(x_pin, non_blocking=True)
            sharding.contract_vec_sharded(RANK, dim, sh_dev, x_dev, begin, end, out, ws)
            return float(out.cpu()[0])
        for _ in range(2):
            e2e_step()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            r_e2e = e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt[0])
        assert abs(r_e2e - result) <= 1e-12 * abs(result), (r_e2e, result)
        e2e = {"value": n_comps / dt, "unit": "packed components/s", "h2d_bytes_per_step": int((end - begin) * 8 + dim * 8),
               "d2h_bytes_per_step": 8, "ms_per_step": dt * 1e3,
               "api": "symtensor_b200.sharding.contract_vec_sharded(pinned host slice per rank -> device, x) + scalar read-back; "
                      "bytes are per rank"}
    if world == 1:
        # ---- CPU baselines on this box's host cores (reported, not a target)
        real = real_reference_arm(1, 0)
        v, dt_cpu, n_cpu = cpu_reference_arm(1, 0)
        cpu_port = {"value": v, "unit": "packed components/s", "cores": os.cpu_count(), "kind": "port",
                    "sample": f"vectorised NumPy port of the reference algorithm (dense + tensordot + r! symmetrize + repack, x4) at dim "
                              f"{CPU_SAMPLE_DIM} ({n_cpu} packed comps, {dt_cpu:.1f} s); dim 200 needs a 12.8 GB dense array"}
        if real is not None:
            rv, rdt, rn, rval = real
            # the GPU path on the same inputs (config 1 through the public API) agrees with the reference's own result
            d1, x1 = synth(REF_CONFIG[0], REF_CONFIG[1], SEED - 1)
            A1 = st.PermClsTorchSymmetricTensor(rank=REF_CONFIG[0], dim=REF_CONFIG[1], data=d1, device=dev)
            g1 = float(st.contract_all_indices_with_vector(A1, x1))
            cpu_baseline = {"value": rv, "unit": "packed components/s", "cores": 1, "kind": "reference", "host_cores": os.cpu_count(),
                            "sample": f"the UNMODIFIED reference (symalg.contract_all_indices_with_vector, PermClsSymmetricTensor, NumPy backend, "
                                      f"one Python thread) on BASELINE configs[0]: rank 4 dim {REF_CONFIG[1]} fp64, {rn} packed comps, one call of "
                                      f"{rdt:.1f} s; configs[1] would take ~37 min and 12.8 GB dense per call",
                            "rel_diff_vs_gpu_same_inputs": abs(g1 - rval) / abs(rval)}
        else:
            cpu_baseline = cpu_port
            cpu_port = None
        try:
            from oracle import c_oracle as co
            host = {c: A._data[c].cpu().numpy() for c in table.classes}
            t0 = time.perf_counter()
            r_c = co.contract_all_indices_with_vector(host, RANK, dim, xh)
            dt_c = time.perf_counter() - t0
            cpu_packed = {"value": n_comps / dt_c, "unit": "packed components/s", "cores": co.num_threads(),
                          "kind": "packed C restatement (OpenMP, long double)", "sample": "full workload",
                          "rel_diff_vs_gpu": abs(r_c - result) / abs(r_c)}
        except Exception as e:  # the oracle is optional for the bench line
            cpu_packed = {"error": str(e)[:200]}

    # ---- the other BASELINE configurations (tensordot, matrix contraction, outer -> vector): extra keys of the line
    other = {}
    if not args.skip_configs:
        import bench_configs as bc
        pk = {"hbm_gbs": measured_peak_hbm()[0], "bf16_tflops": measured_peak_bf16()[0], "measured": measured_peak_bf16()[1]}
        del shard
        torch.cuda.empty_cache()
        for name, fn in (("config5", bc.config5), ("config4", bc.config4), ("config3", bc.config3)):
            try:
                other[name] = fn(dev, world, rank, dist if world > 1 else None, pk)
            except Exception as ex:  # a failed extra must not take the headline line with it
                other[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
            torch.cuda.empty_cache()
        if world == 1:
            try:
                for name, leg in bc.cpu_legs().items():
                    if name in other and "error" not in other[name]:
                        other[name]["cpu_baseline"] = leg
            except Exception as ex:
                other["cpu_legs_error"] = str(ex)[:200]

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        alg_bytes = n_comps * 8 / world  # per launch (per GPU): every stored fp64 value is read exactly once
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "packed components/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"rank 4 dim {dim} float64 permcls contract_all_indices_with_vector"
                                   + (" (BASELINE configs[1])" if world == 1 else f" (configs[1] grown to {world} GPUs, ~6.9e7 comps/GPU)"),
                       "packed_components": n_comps, "packed_bytes": n_comps * 8, "parallelism": f"range-shard x{world}",
                       "collective": "none" if world == 1 else "NCCL all-reduce of 1 fp64 per step, enqueued asynchronously: it overlaps the next step's kernel (two result buffers)",
                       "l2": "input (549 MB per GPU) is larger than the 126 MB L2: no flush needed", "result": result},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(n_comps), "peak_source": peak_src,
                         "kernel": "vec_ring_kernel<double> (one launch per step: per-warp cp.async.bulk rings, dynamic tile deal, per-tile partial sums added in index order by the last CTA; launches chained by programmatic dependent launch -- ST_VEC_OVERLAP -- so that consecutive steps overlap ramp-up and tail)",
                         "algorithmic_bytes_per_launch": alg_bytes},
            "clocks": clocks.summary(),
            "gpu_launches": int(launches),
        }
        line["e2e"] = e2e
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
            if cpu_port is not None:
                line["cpu_port"] = cpu_port
            line["cpu_packed_oracle"] = cpu_packed
        if isolated is not None:
            line["isolated"] = isolated
        if strong is not None:
            line["strong"] = strong
        if serial is not None:
            line["serial"] = serial
        line.update(other)
        line["config"]["launch"] = pipe.how()
        if rebalanced is not None:
            line["config"]["slices"] = "contiguous 32-aligned slices of the packed range, cut points moved until the per-rank kernel times agree"
            line["rebalance"] = rebalanced
        print(json.dumps(line))
    if world > 1:
        # captured NCCL work keeps the communicator busy at teardown (destroy_process_group was seen to hang with live graphs):
        # drop the graphs, drain, and leave without the collective teardown
        pipe.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="N > 1: launch every step from the host instead of replaying a CUDA graph")
    ap.add_argument("--skip-configs", action="store_true", help="only the headline workload (configs[1]); skip configs[2..4]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
