#!/usr/bin/env python
"""One launch of the multiply.outer kernel on a slice of BASELINE config 5 (for ncu): python tools/one_outer.py [fraction]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench_configs as bc  # noqa: E402
from symtensor_b200 import combinatorics as comb, ops  # noqa: E402

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
start = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
if len(sys.argv) > 3:  # python tools/one_outer.py FRACTION START ROWS(0|1)
    from symtensor_b200._cabi import c_i64, check, lib
    check(lib.st_set_tuning(b"outer_rows", c_i64(int(sys.argv[3]))))
dev = torch.device("cuda:0")
A, B = bc.device_tensor(4, 40, 1, torch.float32, dev), bc.device_tensor(4, 40, 2, torch.float32, dev)
af, bf = ops._flat_buffer(A, torch.float32), ops._flat_buffer(B, torch.float32)
total = comb.class_table(8, 40).total
b = int(total * start) // 32 * 32
e = min(total, b + int(total * frac) // 32 * 32)
out = torch.empty(e - b, dtype=torch.float32, device=dev)
for _ in range(2):
    ops.outer_device(A, B, out, b, e, torch.float32, af=af, bf=bf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.outer_device(A, B, out, b, e, torch.float32, af=af, bf=bf)
e1.record()
torch.cuda.synchronize()
print(f"outer r4 (x) r4 dim 40 fp32, coordinates [{b}, {e}): {e0.elapsed_time(e1):.2f} ms for {e - b} components")
