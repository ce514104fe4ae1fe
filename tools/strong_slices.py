#!/usr/bin/env python
"""Per-slice launch times of the vector contraction for the strong-scaling cut of BASELINE config 2 (rank 4 dim 200 fp64 cut
N ways), on ONE GPU, for several tile sizes of the short-launch strategy:  python tools/strong_slices.py [N] [DIM]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb, sharding  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rank = 4
dev = torch.device("cuda:0")
t = comb.class_table(rank, dim)
buf = torch.rand(t.total, dtype=torch.float64, device=dev) + 0.5
x = (torch.rand(dim, dtype=torch.float64, device=dev) + 0.5) / dim ** 0.5
out = torch.zeros(1, dtype=torch.float64, device=dev)
ws = torch.empty(2 * int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)
cuts = sharding.shard_bounds(t.total, n)
whole = None
tmax = int(sys.argv[3]) if len(sys.argv) > 3 else -1
if tmax >= 0:
    check(lib.st_set_tuning(b"vec_ring_table_max", c_i64(tmax)))
    print("vec_ring_table_max", tmax)
for slots in (4,):
    check(lib.st_set_tuning(b"vec_short_launch_bytes", c_i64(0 if slots == 0 else 160 << 20)))
    if slots:
        check(lib.st_set_tuning(b"vec_short_launch_slots", c_i64(slots)))
    times, total = [], 0.0
    for r in range(n):
        b, e = cuts[r], cuts[r + 1]
        for overlap in (0, 1):
            fn = lambda: check(lib.st_contract_vec_ex_f64(0, rank, c_i64(dim), buf[b:].data_ptr(), c_i64(b), c_i64(e), x.data_ptr(), out.data_ptr(),
                                                          ws.data_ptr(), overlap, None))
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 20 * 1e3)
        total += float(out[0])
    iso, ovl = times[0::2], times[1::2]
    print(f"slots {slots or 'default'}: isolated us/slice {[round(v, 1) for v in iso]} max {max(iso):.1f} | overlapped {[round(v, 1) for v in ovl]} max {max(ovl):.1f} | sum of slices {total:.12g}")
