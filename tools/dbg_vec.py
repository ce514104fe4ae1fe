#!/usr/bin/env python
"""Timing of the vector contraction on one class with the kernel's profiling bits (st_set_tuning "vec_debug"):
python tools/dbg_vec.py RANK DIM {f32|f64} CLASS_ORDINAL"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

rank, dim = int(sys.argv[1]), int(sys.argv[2])
tdt = torch.float32 if sys.argv[3] == "f32" else torch.float64
cls = int(sys.argv[4])
dev = torch.device("cuda:0")
t = comb.class_table(rank, dim)
buf = torch.rand(t.total, dtype=tdt, device=dev) + 0.5
x = (torch.rand(dim, dtype=tdt, device=dev) + 0.5) / dim ** 0.5
out = torch.zeros(1, dtype=tdt, device=dev)
ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)
fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32
b, e = (0, t.total) if cls < 0 else (t.offsets[cls], t.offsets[cls + 1])


def run():
    check(fn(0, rank, c_i64(dim), buf[b:].data_ptr(), c_i64(b), c_i64(e), x.data_ptr(), out.data_ptr(), ws.data_ptr(), None))


for tile in (8192, 16384, 32768):
    check(lib.st_set_tuning(b"vec_ring_tile_bytes", c_i64(tile)))
    for slots in (2, 3, 4):
        check(lib.st_set_tuning(b"vec_batch_slots", c_i64(slots)))
        for dbg in (0,):
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print(f"tile={tile} slots={slots} debug={dbg:2d}: {us:8.1f} us  {(e - b) * buf.element_size() / us / 1e3:7.0f} GB/s", flush=True)
