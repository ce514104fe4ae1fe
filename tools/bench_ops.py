#!/usr/bin/env python
"""Timing of the tensor-valued ops at BASELINE.json's configs (3, 4, 5) or the largest size one GPU holds:
python tools/bench_ops.py [outer|mat|tdot|all]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import symtensor_b200 as st  # noqa: E402
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")


def rand_tensor(rank, dim, tdt, seed):
    t = comb.class_table(rank, dim)
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    buf = torch.rand(t.total, generator=g, dtype=tdt, device=DEV) + 0.5
    for c, s, o in zip(t.classes, t.sizes, t.offsets):
        buf[o + s:t.offsets[t.index(c) + 1]] = 0
    return st.PermClsTorchSymmetricTensor.from_packed(rank, dim, buf)


def timeit(f, reps=3, warm=1):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


def bench_outer():
    # config 5: rank 8 dim 40 fp32 = outer of two rank-4 tensors, then the vector contraction
    A, B = rand_tensor(4, 40, torch.float32, 1), rand_tensor(4, 40, torch.float32, 2)
    x = ((torch.rand(40, dtype=torch.float32) + 0.5) / 40 ** 0.5).numpy()
    n = comb.indep_size(8, 40)
    ms, C = timeit(lambda: st.multiply.outer(A, B), reps=2)
    print(f"outer r4 (x) r4 dim 40 fp32 -> rank 8: {ms:9.2f} ms  {n / ms / 1e6:8.2f} G comps/s  {n * 4 / ms / 1e6:7.1f} GB/s written")
    ms2, s = timeit(lambda: st.contract_all_indices_with_vector(C, x), reps=5)
    print(f"  then contract_all_indices_with_vector:  {ms2:9.3f} ms  {n * 4 / ms2 / 1e6:7.1f} GB/s   value {float(s):.6g}")
    ms3, s3 = timeit(lambda: ops.outer_then_contract_vec(A, B, x), reps=2)
    ident = float(st.contract_all_indices_with_vector(A, x)) * float(st.contract_all_indices_with_vector(B, x))
    print(f"  fused outer->vector (nothing stored):  {ms3:9.2f} ms   value {float(s3):.6g}  identity (A.x^4)(B.x^4) = {ident:.6g}")


def bench_mat():
    # config 4: rank 6 dim 64 fp64 with a 64 x 64 W
    rank, dim = 6, 64
    A = rand_tensor(rank, dim, torch.float64, 3)
    rng = np.random.default_rng(4)
    W = rng.uniform(0.5, 1.5, (dim, dim)) / dim
    y = rng.uniform(0.5, 1.5, dim)
    n = comb.indep_size(rank, dim)
    from oracle import packed_oracle as po
    flops = po.semi_packed_chain_flops(rank, dim)
    ms, C = timeit(lambda: st.contract_all_indices_with_matrix(A, W), reps=2)
    lhs = float(st.contract_all_indices_with_vector(C, y))
    rhs = float(st.contract_all_indices_with_vector(A, W @ y))
    print(f"contract_all_indices_with_matrix r6 d64 fp64: {ms:9.1f} ms  {n / ms / 1e3:8.2f} M comps/s  {flops / ms / 1e9:6.2f} TFLOP/s (algorithmic "
          f"{flops:.3e} flops)  identity rel err {abs(lhs - rhs) / abs(rhs):.2e}")


def bench_tdot():
    # config 3 is rank 3 dim 1000 fp32, k = 1 (output 167 GB): beyond one GPU; largest materialised-Gram size here
    for ra, rb, k, dim in [(3, 3, 1, 160), (3, 2, 1, 400), (3, 2, 1, 1000)]:  # the last: SURVEY.md 8d fallback ladder for config 3
        A, B = rand_tensor(ra, dim, torch.float32, 5), rand_tensor(rb, dim, torch.float32, 6)
        n = comb.indep_size(ra + rb - 2 * k, dim)
        flops = 2 * dim * comb.indep_size(ra - k, dim) * comb.indep_size(rb - k, dim)
        ms, C = timeit(lambda: st.tensordot(A, B, axes=k), reps=2)
        print(f"tensordot r{ra}.r{rb} k={k} dim {dim} fp32 -> rank {ra + rb - 2 * k}: {ms:9.2f} ms  {n / ms / 1e6:8.3f} G comps/s  Gram "
              f"{flops / ms / 1e9:6.2f} TFLOP/s")


def bench_pack():
    # dense <-> packed: HBM-bound on the dense side, dim^rank elements written / read once
    for rank, dim, tdt in [(4, 128, torch.float64), (3, 800, torch.float32), (6, 24, torch.float64)]:
        A = rand_tensor(rank, dim, tdt, 7)
        es = A.packed.element_size()
        nd = dim ** rank
        ms, dense = timeit(lambda: A.todense(), reps=5)
        print(f"todense r{rank} d{dim} {str(tdt)[6:]}: {ms:8.3f} ms  {nd * es / ms / 1e6:7.1f} GB/s of dense bytes written ({nd * es / 1e9:.2f} GB)")
        ms2, B = timeit(lambda: st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=dense, device=DEV), reps=5)
        print(f"  pack + symmetry check:      {ms2:8.3f} ms  {nd * es / ms2 / 1e6:7.1f} GB/s of dense bytes read   round trip exact: "
              f"{bool(torch.equal(B.packed, A.packed))}")
        ms3, _ = timeit(lambda: st.PermClsTorchSymmetricTensor(rank=rank, dim=dim, data=dense, symmetrize=True, device=DEV), reps=5)
        print(f"  pack with symmetrize:       {ms3:8.3f} ms  {nd * es / ms3 / 1e6:7.1f} GB/s")
    A, B = rand_tensor(3, 100, torch.float64, 8), rand_tensor(3, 100, torch.float64, 9)
    n = comb.indep_size(6, 100)
    ms, _ = timeit(lambda: st.symalg.add.outer(A, B), reps=2)
    print(f"add.outer r3 (+) r3 dim 100 fp64 -> rank 6 ({n} comps): {ms:8.2f} ms  {n / ms / 1e6:7.3f} G comps/s")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("pack", "all"):
        bench_pack()
    if what in ("outer", "all"):
        bench_outer()
    if what in ("mat", "all"):
        bench_mat()
    if what in ("tdot", "all"):
        bench_tdot()
