#!/usr/bin/env python
"""Sweep the ring geometry of vec_ring_kernel on a B200 (run under gpurun).

    python tools/tune_ring.py [RANK DIM {f32|f64}] [--old]
One line per (warps, slots, slot bytes, tile bytes): us per launch, achieved GB/s of algorithmic bytes, value.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

DEV = torch.device("cuda:0")


def bench(rank, dim, tdt, reps=30):
    t = comb.class_table(rank, dim)
    torch.manual_seed(0)
    buf = torch.rand(t.total, dtype=tdt, device=DEV) + 0.5
    x = (torch.rand(dim, dtype=tdt, device=DEV) + 0.5) / dim ** 0.5
    out = torch.zeros(1, dtype=tdt, device=DEV)
    ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=DEV)
    fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32

    def run():
        check(fn(0, rank, c_i64(dim), buf.data_ptr(), c_i64(0), c_i64(t.total), x.data_ptr(), out.data_ptr(), ws.data_ptr(), None))
    for _ in range(3):
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
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rank, dim = (int(args[0]), int(args[1])) if len(args) >= 2 else (4, 200)
    tdt = torch.float32 if (len(args) >= 3 and args[2] == "f32") else torch.float64
    if "--old" in sys.argv:
        check(lib.st_set_vec_variant(3))
        ms, gbs, val = bench(rank, dim, tdt)
        print(f"r{rank} d{dim} {str(tdt)[6:]} OLD tail kernel: {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  val={val:.12g}", flush=True)
        check(lib.st_set_vec_variant(0))
    combos = []
    for warps, slots, nbytes, tmax in [(16, 2, 4096, 64), (16, 2, 4096, 56), (16, 2, 4096, 72), (16, 2, 4096, 48), (16, 2, 3584, 80), (16, 2, 3072, 96),
                                       (16, 3, 2560, 64), (14, 2, 4096, 80), (12, 2, 4096, 96), (16, 2, 4608, 48), (16, 2, 5120, 40)]:
        combos.append((warps, slots, nbytes, 32768, tmax))
    for tile in (16384, 24576, 49152, 65536):
        combos.append((16, 2, 4096, tile, 64))
    if "--static" in sys.argv:
        check(lib.st_set_tuning(b"vec_ring_dynamic", c_i64(0)))
    for warps, slots, nbytes, tile, tmax in combos:
        check(lib.st_set_tuning(b"vec_ring_table_max", c_i64(tmax * 1024)))
        check(lib.st_set_tuning(b"vec_ring_warps", c_i64(warps)))
        check(lib.st_set_tuning(b"vec_ring_slots", c_i64(slots)))
        check(lib.st_set_tuning(b"vec_ring_bytes", c_i64(nbytes)))
        check(lib.st_set_tuning(b"vec_ring_bytes_max", c_i64(nbytes)))
        check(lib.st_set_tuning(b"vec_ring_tile_bytes", c_i64(tile)))
        try:
            ms, gbs, val = bench(rank, dim, tdt)
            print(f"r{rank} d{dim} {str(tdt)[6:]} warps={warps:2d} slots={slots} bytes={nbytes:5d} tile={tile:6d} tmax={tmax:3d}K: {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  val={val:.12g}",
                  flush=True)
        except Exception as exc:  # noqa: BLE001
            print(f"r{rank} d{dim} warps={warps} slots={slots} bytes={nbytes} tile={tile}: FAILED {exc}", flush=True)


if __name__ == "__main__":
    main()
