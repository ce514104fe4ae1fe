#!/usr/bin/env python
"""A few launches of the vector contraction for profiling: python tools/one_vec.py RANK DIM {f32|f64} [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

rank, dim = int(sys.argv[1]), int(sys.argv[2])
tdt = torch.float32 if sys.argv[3] == "f32" else torch.float64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
only_cls = int(sys.argv[5]) if len(sys.argv) > 5 else -1  # restrict the range to one class ordinal
dev = torch.device("cuda:0")
t = comb.class_table(rank, dim)
buf = torch.rand(t.total, dtype=tdt, device=dev) + 0.5
x = (torch.rand(dim, dtype=tdt, device=dev) + 0.5) / dim ** 0.5
out = torch.zeros(1, dtype=tdt, device=dev)
ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)
fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32
b, e = (0, t.total) if only_cls < 0 else (t.offsets[only_cls], t.offsets[only_cls + 1])
for _ in range(reps):
    check(fn(0, rank, c_i64(dim), buf[b:].data_ptr(), c_i64(b), c_i64(e), x.data_ptr(), out.data_ptr(), ws.data_ptr(), None))
torch.cuda.synchronize()
print("value", float(out[0]))
