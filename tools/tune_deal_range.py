#!/usr/bin/env python
"""Guided deal on a SUB-RANGE of the packed tensor (what one rank of a sharded run walks):
    python tools/tune_deal_range.py RANK DIM PARTS PART   # times part PART of PARTS equal slices with group = 1 and 4"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

rank, dim, parts, part = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda:0")
t = comb.class_table(rank, dim)
buf = torch.rand(t.total, dtype=torch.float64, device=dev) + 0.5
x = (torch.rand(dim, dtype=torch.float64, device=dev) + 0.5) / dim ** 0.5
out = torch.zeros(1, dtype=torch.float64, device=dev)
ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)
b = (t.total * part // parts) // 32 * 32
e = t.total if part == parts - 1 else (t.total * (part + 1) // parts) // 32 * 32
for group in (1, 4, 2, 3):
    check(lib.st_set_tuning(b"vec_ring_group", c_i64(group)))
    f = lambda: check(lib.st_contract_vec_f64(0, rank, c_i64(dim), buf[b:].data_ptr(), c_i64(b), c_i64(e), x.data_ptr(), out.data_ptr(), ws.data_ptr(), None))  # noqa: E731
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"r{rank} d{dim} part {part}/{parts} [{b}, {e}) group={group}: {us:8.1f} us  {(e - b) * 8 / us / 1e3:7.0f} GB/s  val={float(out[0]):.12g}", flush=True)
