#!/usr/bin/env python
"""The row walk on the device against its host replay (both against the oracle in tests): python tools/rowwalk_device_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

for rank, dim, span in [(1, 7, 2048), (2, 9, 2048), (3, 6, 64), (4, 11, 2048), (5, 5, 96), (6, 7, 2048), (8, 5, 2048), (8, 9, 2048), (4, 40, 1024)]:
    total = comb.class_table(rank, dim).total
    host = np.full((total, rank), -7, dtype=np.int32)
    assert lib.st_debug_rowwalk(rank, c_i64(dim), c_i64(0), c_i64(total), c_i64(span), host.ctypes.data) == total
    dev = torch.full((total, rank), -9, dtype=torch.int32, device="cuda")
    check(lib.st_debug_rowwalk_device(rank, c_i64(dim), c_i64(0), c_i64(total), c_i64(span), dev.data_ptr(), None))
    torch.cuda.synchronize()
    got = dev.cpu().numpy()
    bad = np.nonzero((got != host).any(axis=1))[0]
    print(rank, dim, span, "ok" if len(bad) == 0 else f"{len(bad)} of {total} differ; first {bad[:6]} device {got[bad[:3]].tolist()} host {host[bad[:3]].tolist()}")
