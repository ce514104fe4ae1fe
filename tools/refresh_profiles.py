#!/usr/bin/env python
"""Turn the artefacts a validation run left in gpurun_out/ into the tracked summaries under profiles/:
    python tools/refresh_profiles.py
(ring_prof_final.ncu-rep -> r1_vec_c2_ncu_full.txt / r1_vec_c2_hot_lines.txt / roofline_traffic.json, launches.csv ->
r1_launches_bench.csv / _summary.txt, bench_r1.json -> r1_bench_n1.json, tl.log, bench_ops.log)."""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = os.path.join(G, "ring_prof_final.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    vals = {w: (r[hdr.index(w)], units[hdr.index(w)]) for w in WANT if w in hdr}
    with open(os.path.join(P, "r1_vec_c2_ncu_full.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on -k regex:vec_ring -s 1 -c 1  python tools/one_vec.py 4 200 f64 3\n"
                "(rank 4 dim 200 fp64 whole tensor, 68,685,050 packed components = 549,480,400 algorithmic bytes; times under ncu are "
                "cold-cache and serialised)\n\n")
        for w, (v, u) in vals.items():
            f.write(f"{w:90s} {v} {u}\n")
    hot = subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "25"], capture_output=True, text=True).stdout
    with open(os.path.join(P, "r1_vec_c2_hot_lines.txt"), "w") as f:
        f.write("tools/ncu_hot.py on the same report: share of stall samples / executed warp-instructions per CUDA source line\n" + hot)

    def mbytes(key):
        v, u = vals[key]
        return float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u]
    rd, wr = mbytes("dram__bytes_read.sum"), mbytes("dram__bytes_write.sum")
    json.dump({"packed_components": 68685050, "dram_bytes_per_launch": rd + wr,
               "source": f"profiles/r1_vec_c2_ncu_full.txt (ncu --set full, vec_ring_kernel<double>, rank 4 dim 200 fp64 whole tensor; "
                         f"read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB)"}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, "r1_launches_bench.csv"))
    rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
    hi = [i for i, x in enumerate(rows) if x and x[0] == "ID"][0]
    h = rows[hi]
    ki, vi, gi, bi, ui = (h.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size", "Metric Unit"))
    agg = collections.OrderedDict()
    for x in rows[hi + 1:]:
        if len(x) < len(h):
            continue
        v = float(x[vi].replace(",", "")) / (1e3 if x[ui] in ("ns", "nsecond") else 1.0)
        a = agg.setdefault((x[ki][:92], x[gi], x[bi]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "r1_launches_bench_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 60  python bench.py --steps 3 --warmup 3   (first 60 launches: "
                "cold-cache, serialised -- compare shares; grids below 148 CTAs are the chunks of the host-streamed e2e leg)\n")
        f.write(f"{'kernel':92s} {'launches':>8s} {'total us':>10s} {'share':>7s}  grid block\n")
        for (k, g, b), (n, t) in agg.items():
            f.write(f"{k:92s} {n:8d} {t:10.1f} {100 * t / tot:6.1f}%  {g} {b}\n")
    shutil.copy(os.path.join(G, "bench_r1.json"), os.path.join(P, "r1_bench_n1.json"))
    shutil.copy(os.path.join(G, "tl.log"), os.path.join(P, "r1_vec_c2_timeline.txt"))
    if os.path.exists(os.path.join(G, "bench_ops.log")):
        shutil.copy(os.path.join(G, "bench_ops.log"), os.path.join(P, "r1_bench_ops.txt"))
    j = json.load(open(os.path.join(P, "r1_bench_n1.json")))
    print("bench:", j["ms_per_step"] * 1e3, "us", j["roofline"]["frac"], "traffic", rd + wr)


if __name__ == "__main__":
    main()
