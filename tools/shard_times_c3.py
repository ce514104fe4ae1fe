#!/usr/bin/env python
"""Time of every GPU's shard of BASELINE config 3 (sharding.tensordot22_shards, five ranges in one call) on ONE GPU, with the
number of tiles each runs:  python tools/shard_times_c3.py [WORLD]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench_configs as bc  # noqa: E402
from symtensor_b200 import ops, sharding  # noqa: E402
from symtensor_b200._cabi import c_i64, lib  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dim = 1000
dev = torch.device("cuda:0")
A, B = bc.device_tensor(3, dim, 1, torch.float32, dev), bc.device_tensor(3, dim, 2, torch.float32, dev)
af, bf = ops._flat_buffer(A, torch.float32), ops._flat_buffer(B, torch.float32)
shards = sharding.tensordot22_shards(dim, world)
ws = None
for g, ranges in enumerate(shards):
    n = len(ranges)
    bs, es = (ctypes.c_int64 * n)(*[r[0] for r in ranges]), (ctypes.c_int64 * n)(*[r[1] for r in ranges])
    nt = int(lib.st_debug_sym22_tiles_ranges(c_i64(dim), n, bs, es, ctypes.c_void_p(0), c_i64(0)))
    outs = [torch.empty(e - b, dtype=torch.float32, device=dev) for b, e in ranges]
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ws = ops.tensordot_device_ranges(A, B, 1, outs, ranges, af=af, bf=bf, ws=ws, check_flag=False)
        e1.record()
        torch.cuda.synchronize()
    print(f"shard {g}: {nt} tiles, {sum(e - b for b, e in ranges)} comps, {e0.elapsed_time(e1):.1f} ms ({e0.elapsed_time(e1) / nt * 1e3:.3f} us per tile)", flush=True)
    del outs
