#!/usr/bin/env python
"""Sweep the tuning knobs of the vector-contraction kernel on a B200 (run under gpurun).

    python tools/tune_vec.py [--quick]
Prints one line per (workload, threads, tile_bytes): ms per launch and achieved GB/s of algorithmic bytes.
"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

DEV = torch.device("cuda:0")


def bench(rank, dim, tdt, reps=50):
    t = comb.class_table(rank, dim)
    buf = torch.rand(t.total, dtype=tdt, device=DEV) + 0.5
    x = (torch.rand(dim, dtype=tdt, device=DEV) + 0.5) / dim ** 0.5
    out = torch.zeros(1, dtype=tdt, device=DEV)
    ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=DEV)
    fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32

    def run():
        check(fn(0, rank, c_i64(dim), buf.data_ptr(), c_i64(0), c_i64(t.total), x.data_ptr(), out.data_ptr(), ws.data_ptr(), None))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = sum(t.sizes) * buf.element_size()
    return ms, nbytes / ms / 1e6, float(out[0])


def main():
    quick = "--quick" in sys.argv
    workloads = [(4, 200, torch.float64), (8, 40, torch.float32), (6, 64, torch.float64), (4, 200, torch.float32), (3, 1000, torch.float32)]
    if quick:
        workloads = workloads[:2]
    for rank, dim, tdt in workloads:
        for variant in ([0] if quick else [0]):
            for ipc in [4096, 8192, 16384, 32768, 65536]:
                check(lib.st_set_tuning(b"vec_tile_bytes", c_i64(ipc)))
                ms, gbs, val = bench(rank, dim, tdt)
                print(f"r{rank} d{dim} {str(tdt)[6:]} tile_bytes={ipc}: {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  val={val:.6g}", flush=True)


if __name__ == "__main__":
    main()
