#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS stall samples / instruction counts by CUDA source line.
    python tools/ncu_hot.py REPORT.ncu-rep [topN]
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    agg = defaultdict(lambda: [0, 0, ""])
    fpath, hdr = None, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            ns, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        key = (fpath, line)
        try:
            agg[key][0] += int(float(r[ns] or 0))
            agg[key][1] += int(float(r[ie] or 0))
        except ValueError:
            pass
        if r[1].strip():
            agg[key][2] = r[1].strip()
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {ts}  warp-instructions {ti}")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / ts:5.1f}% smp {100 * v[1] / ti:5.1f}% ins  {f}:{l:<4d} {v[2][:100]}")


if __name__ == "__main__":
    main()
