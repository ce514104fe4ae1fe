#!/usr/bin/env python
"""Timing of the tiled tensordot kernel (st_sym22.cu) on BASELINE config 3's operands: one GPU computes the output range an
8-GPU run gives it (shares of the packed rank-4 dim-1000 output).  CUDA events around the C-ABI call; expand + memset +
kernel are reported together and the kernel alone (second call, operands already expanded is not possible through the
C-ABI, so the expand kernels are timed separately by a dim^2 Kp copy model -- see the ncu launch list for the split).

    python tools/bench_sym22.py [--dim 1000] [--world 8] [--shares 0,4] [--kch 16,32]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import symtensor_b200 as st  # noqa: E402
from symtensor_b200 import combinatorics as comb, ops, sharding  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402


def device_tensor(rank, dim, seed, dtype):
    t = comb.class_table(rank, dim)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    buf = (torch.rand(t.total, generator=g, dtype=torch.float64, device="cuda") + 0.5).to(dtype)
    for i in range(t.ncls):
        buf[t.offsets[i] + t.sizes[i]:t.offsets[i + 1]] = 0
    return st.PermClsTorchSymmetricTensor.from_packed(rank, dim, buf)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=1000)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--shares", default="0,4")
    ap.add_argument("--kch", default="16,32")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--rgroup", default="", help="values of the tile-order key sym22_rgroup to time (default: the library's)")
    ap.add_argument("--balanced", type=int, default=0, help="1: the tile-balanced cuts of sharding.tensordot22_bounds (what bench.py shards with)")
    ap.add_argument("--debug", default="0", help="ablation masks to time (sym22_debug): 1 no adds, 2 no drains, 4 no TMA, 8 no MMA")
    args = ap.parse_args()
    dim = args.dim
    A, B = device_tensor(3, dim, 1, torch.float32), device_tensor(3, dim, 2, torch.float32)
    af, bf = ops._flat_buffer(A, torch.float32), ops._flat_buffer(B, torch.float32)
    table = comb.class_table(4, dim)
    cuts = sharding.tensordot22_bounds(dim, args.world) if args.balanced else sharding.shard_bounds(table.total, args.world)
    ws = None
    for rg, dbg, kch in [(r, int(g), int(v)) for r in (args.rgroup.split(",") if args.rgroup else [""]) for g in args.debug.split(",")
                         for v in args.kch.split(",")]:
        if rg:
            check(lib.st_set_tuning(b"sym22_rgroup", c_i64(int(rg))))
            print("tile order: groups of", rg, "k blocks")
        check(lib.st_set_tuning(b"sym22_kch", c_i64(kch)))
        check(lib.st_set_tuning(b"sym22_debug", c_i64(dbg)))
        print("ablation mask", dbg)
        for sh in [int(v) for v in args.shares.split(",")]:
            b, e = cuts[sh], cuts[sh + 1]
            out = torch.empty(e - b, dtype=torch.float32, device="cuda")
            for rep in range(args.reps):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ws = ops.tensordot_device(A, B, 1, out, b, e, torch.float32, af=af, bf=bf, ws=ws, check_flag=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                flag = int(ws[:1].view(torch.int32)[0].item())
                if dbg & 64:
                    c = ws[:64].view(torch.int64)[8:14].cpu().tolist()
                    nw = max(c[5], 1)
                    print("   epilogue warp time (M clocks per warp): wait %.1f drain %.1f write %.1f fence %.1f barrier %.1f  (warps %d)" %
                          tuple([v / nw / 1e6 for v in c[:5]] + [nw]))
                n = e - b
                flops = 12.0 * dim * n  # 2 d C(4,2) per packed output component
                print(f"dim {dim} kch {kch} share {sh}/{args.world}: {n} comps ({n * 4 / 1e9:.2f} GB) in {ms:.1f} ms = "
                      f"{n / ms / 1e6:.2f} G comps/s, {flops / ms / 1e9:.1f} TFLOP/s fp32-equivalent, {3 * flops / ms / 1e9:.1f} TFLOP/s on the tensor pipe "
                      f"(3xTF32), err flag {flag}", flush=True)
            del out


if __name__ == "__main__":
    main()
