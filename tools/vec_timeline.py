#!/usr/bin/env python
"""Phase timeline of one vec_ring_kernel launch (globaltimer stamps written by the kernel in debug mode).

    python tools/vec_timeline.py RANK DIM {f32|f64} [class ordinal]
Prints, relative to the earliest CTA start: per phase the min / median / max over CTAs, and the spread of the
per-warp finish times.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symtensor_b200 import combinatorics as comb  # noqa: E402
from symtensor_b200._cabi import c_i64, check, lib  # noqa: E402

rank, dim = int(sys.argv[1]), int(sys.argv[2])
tdt = torch.float32 if sys.argv[3] == "f32" else torch.float64
only_cls = int(sys.argv[4]) if len(sys.argv) > 4 else -1
dev = torch.device("cuda:0")
t = comb.class_table(rank, dim)
buf = torch.rand(t.total, dtype=tdt, device=dev) + 0.5
x = (torch.rand(dim, dtype=tdt, device=dev) + 0.5) / dim ** 0.5
out = torch.zeros(1, dtype=tdt, device=dev)
ws = torch.empty(int(lib.st_contract_vec_workspace_bytes()) // 8, dtype=torch.float64, device=dev)
fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32
b, e = (0, t.total) if only_cls < 0 else (t.offsets[only_cls], t.offsets[only_cls + 1])


def run():
    check(fn(0, rank, c_i64(dim), buf[b:].data_ptr(), c_i64(b), c_i64(e), x.data_ptr(), out.data_ptr(), ws.data_ptr(), None))


for _ in range(3):
    run()
torch.cuda.synchronize()
check(lib.st_set_tuning(b"vec_timeline", c_i64(1)))
run()
torch.cuda.synchronize()
n = 148 * 16 + 4096 + 4096 * 8
raw = np.zeros(n, dtype=np.uint64)
check(lib.st_debug_vec_timeline(raw.ctypes.data, c_i64(n)))
ph = raw[:148 * 16].reshape(148, 16).astype(np.int64)
live = ph[:, 0] > 0
ph = ph[live]
t0 = ph[:, 0].min()
names = ["start", "cls+bars+runs", "stream started", "x/binom/cdesc+table", "small classes done (warp 0)"] + \
        ["class 0 begins", "last CTA: reduce_slots done", "class 2 begins", "class 3 begins", "class 4 begins"] + \
        ["all warps of the CTA done", "ticket taken", "last CTA: counters reset", "last CTA: ranges built", "last CTA: slots added", "end"]
print(f"{live.sum()} CTAs; times in us after the first CTA start")
for s, name in enumerate(names):
    col = ph[:, s]
    col = col[col > 0]
    if len(col) == 0:
        continue
    r = (col - t0) / 1e3
    print(f"  [{s:2d}] {name:30s} min {r.min():8.2f}  med {np.median(r):8.2f}  max {r.max():8.2f}")
wf_all = raw[148 * 16:148 * 16 + 4096].astype(np.int64)
wf = wf_all[wf_all > 0]
if len(wf):
    r = (wf - t0) / 1e3
    q = np.percentile(r, [0, 5, 25, 50, 75, 95, 100])
    print("  warp finish times: " + "  ".join(f"p{p}={v:.1f}" for p, v in zip([0, 5, 25, 50, 75, 95, 100], q)))

ce = raw[148 * 16 + 4096:].reshape(4096, 8).astype(np.int64)
for ci in range(5):
    col = ce[:, ci]
    m = col > 0
    if m.sum() == 0:
        continue
    r = (col[m] - t0) / 1e3
    q = np.percentile(r, [0, 50, 95, 99, 100])
    print(f"  warps leave class {ci}: " + "  ".join(f"p{p}={v:.1f}" for p, v in zip([0, 50, 95, 99, 100], q)))
late = np.argsort(-wf_all)[:12]
print("  latest warps (cta, warp, finish us, left class 3 at us, tiles walked, longest tile us (directory index), last tile began at us):")
for w in late:
    if wf_all[w] > 0:
        print(f"    cta {w // 16:3d} warp {w % 16:2d}  {(wf_all[w] - t0) / 1e3:7.1f}  {(ce[w, 3] - t0) / 1e3 if ce[w, 3] > 0 else -1:7.1f}  {ce[w, 5]:3d}  {ce[w, 6] / 1e3:6.1f} ({ce[w, 2]})  {(ce[w, 7] - t0) / 1e3:7.1f}")
m = wf_all > 0
print(f"  tiles per warp: min {ce[m, 5].min()} med {np.median(ce[m, 5]):.0f} max {ce[m, 5].max()};  longest tile over all warps {ce[m, 6].max() / 1e3:.1f} us, median of the per-warp longest {np.median(ce[m, 6]) / 1e3:.1f} us")
