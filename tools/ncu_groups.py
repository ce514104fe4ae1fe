#!/usr/bin/env python
"""Group an ncu report's per-line instruction counts / samples of vec_ring_kernel by code region.
    python tools/ncu_groups.py REPORT.ncu-rep
"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    agg = defaultdict(lambda: [0, 0])
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
            agg[(fpath, line)][0] += int(float(r[ns] or 0))
            agg[(fpath, line)][1] += int(float(r[ie] or 0))
        except ValueError:
            pass
    return agg


def markers(path, pats):
    src = open(path).read().split("\n")
    out = []
    for name, pat in pats:
        for i, l in enumerate(src):
            if pat in l:
                out.append((i + 1, name))
                break
        else:
            raise KeyError(pat)
    return sorted(out)


def main():
    agg = load(sys.argv[1])
    core = markers(os.path.join(ROOT, "symtensor_b200/csrc/st_vec_core.cuh"), [
        ("core: helpers (unrank_earlier, xrel_pow, comb_unrank_warp)", "ST_HD double unrank_earlier("),
        ("walk: init + advance", "ST_HD double walk_tile(const PlanView& P"),
        ("walk: row walk", "auto row_range = [&]"),
        ("walk: sub-chunk / piece control", "const int Bel = src.bel();"),
        ("walk: table loop", "const T* __restrict__ tp = tbl + (toff - qs);"),
        ("walk: flush / advance control", "if (pe > ce) break;  // the piece continues"),
        ("walk: dispatch", "ST_HD double walk_tile_any("),
        ("core: dir_decode / table builds", "ST_HD double dir_decode("),
        ("core: make_runs", "ST_HD void make_runs("),
        ("ring: ptx helpers (wait / copy asm)", "__device__ __forceinline__ uint32_t smem_u32"),
        ("ring: claims + set_tile + enter_class + next_tile + start", "struct RingSrc {"),
        ("ring: wait_slot + issue", "ST_HD void wait_slot() {"),
        ("ring: head_is / pop / open_tile", "ST_HD bool head_is("),
        ("ring: release / chunk / close_tile", "ST_HD void release() {"),
        ("core: layout", "struct RingLayout {"),
    ])
    kern = markers(os.path.join(ROOT, "symtensor_b200/csrc/st_vec.cu"), [
        ("kernel: (other kernels)", "__global__ void vec_finalize_kernel"),
        ("kernel: prologue", "vec_ring_kernel(const __grid_constant__"),
        ("kernel: prologue table build", "// the first class with a class-wide table"),
        ("kernel: small classes", "// ---- 4. SMALL classes"),
        ("kernel: class loop setup", "// ---- 3. tail-table classes, in stream order"),
        ("kernel: mode A tile loop", "while (src.head_is(ci)) {"),
        ("kernel: mode B", "// ---- mode B: chunks of nwarps tiles per CTA (static deal)"),
        ("kernel: epilogue (ticket, final reduce)", "// one partial per warp of the grid, added in index order by the last CTA"),
    ])

    def grp(f, l):
        tab = core if f == "st_vec_core.cuh" else kern if f == "st_vec.cu" else None
        if tab is None:
            return "other: " + str(f)
        name = "before first marker (" + f + ")"
        for ln, nm in tab:
            if l >= ln:
                name = nm
        return name
    g = defaultdict(lambda: [0, 0])
    for (f, l), v in agg.items():
        k = grp(f, l)
        g[k][0] += v[0]
        g[k][1] += v[1]
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {ts}  warp-instructions {ti / 1e6:.1f}M")
    for k, v in sorted(g.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * v[1] / ti:5.1f}% ins ({v[1] / 1e6:5.1f}M) {100 * v[0] / ts:5.1f}% smp  {k}")


if __name__ == "__main__":
    main()
