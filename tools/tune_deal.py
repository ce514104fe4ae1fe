#!/usr/bin/env python
"""Sweep the dynamic deal of vec_ring_kernel on a B200 (run under gpurun): tile size (directory granularity), tiles
per group at the start of a class, single tiles per warp at its end.

    python tools/tune_deal.py [RANK DIM {f32|f64}]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tune_ring import bench  # noqa: E402
import torch  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rank, dim = (int(args[0]), int(args[1])) if len(args) >= 2 else (4, 200)
    tdt = torch.float32 if (len(args) >= 3 and args[2] == "f32") else torch.float64
    combos = [(49152, 1, 0, 2), (49152, 3, 2, 0)]
    for tile in (49152, 32768, 24576, 16384):
        for fine in (1, 2, 3, 4):
            combos.append((tile, 0, fine, 0))
    combos += [(49152, 4, 1, 0), (49152, 4, 2, 0), (65536, 2, 1, 0), (65536, 2, 2, 0), (65536, 3, 1, 0), (32768, 4, 3, 0), (32768, 5, 2, 0)]
    for a in sys.argv[1:]:
        if a.startswith("--combos="):  # tile:group:fine:ondemand,...
            combos = [tuple(int(v) for v in c.split(":")) for c in a[len("--combos="):].split(",")]
    if "--short" in sys.argv:
        combos = [c for c in combos if c[0] in (49152, 24576)]
    combos = [c if len(c) > 4 else tuple(c) + (60,) for c in combos]
    for tile, group, fine, od, pos in combos:
        check(lib.st_set_tuning(b"vec_ring_fine_pos", c_i64(pos)))
        check(lib.st_set_tuning(b"vec_ring_tile_bytes", c_i64(tile)))
        check(lib.st_set_tuning(b"vec_ring_group", c_i64(group)))
        check(lib.st_set_tuning(b"vec_ring_fine", c_i64(fine)))
        check(lib.st_set_tuning(b"vec_ring_ondemand", c_i64(od)))
        try:
            ms, gbs, val = bench(rank, dim, tdt, reps=50)
            print(f"r{rank} d{dim} {str(tdt)[6:]} tile={tile:6d} group={group:2d} fine={fine:2d} ondemand={od} pos={pos:3d}: {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  val={val:.15g}", flush=True)
        except Exception as exc:  # noqa: BLE001
            print(f"r{rank} d{dim} tile={tile} group={group} fine={fine}: FAILED {exc}", flush=True)


if __name__ == "__main__":
    main()
