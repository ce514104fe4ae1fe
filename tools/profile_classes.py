#!/usr/bin/env python
"""Per-class timing of the vector-contraction kernel (uses the [begin, end) range interface to run one class at
a time).  python tools/profile_classes.py RANK DIM {f32|f64} [tile_bytes]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

DEV = torch.device("cuda:0")


def main():
    rank, dim = int(sys.argv[1]), int(sys.argv[2])
    tdt = torch.float32 if sys.argv[3] == "f32" else torch.float64
    if len(sys.argv) > 4:
        check(lib.st_set_tuning(b"vec_ring_tile_bytes", c_i64(int(sys.argv[4]))))
    t = comb.class_table(rank, dim)
    buf = torch.rand(t.total, dtype=tdt, device=DEV) + 0.5
    x = (torch.rand(dim, dtype=tdt, device=DEV) + 0.5) / dim ** 0.5
    out = torch.zeros(1, dtype=tdt, device=DEV)
    ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=DEV)
    fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32
    es = buf.element_size()

    def timeit(b, e, reps=20):
        def run():
            check(fn(0, rank, c_i64(dim), buf[b:].data_ptr(), c_i64(b), c_i64(e), x.data_ptr(), out.data_ptr(), ws.data_ptr(), None))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    tot = 0.0
    for c, (cls, size) in enumerate(zip(t.classes, t.sizes)):
        if size == 0:
            continue
        ms = timeit(t.offsets[c], t.offsets[c + 1])
        tot += ms
        print(f"class {str(cls):28s} size {size:12d}  {ms * 1e3:9.1f} us  {size * es / ms / 1e6:8.0f} GB/s", flush=True)
    ms = timeit(0, t.total)
    print(f"whole tensor: {ms * 1e3:.1f} us  {sum(t.sizes) * es / ms / 1e6:.0f} GB/s   (sum of classes {tot * 1e3:.1f} us)")


if __name__ == "__main__":
    main()
