#!/usr/bin/env python
"""Text summary of one kernel of an ncu report for profiles/:  python tools/ncu_summary.py REPORT.ncu-rep OUT.txt "how it was captured"
(selected raw metrics, the stall reasons per issue, and the hottest CUDA source lines by samples / executed instructions)."""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    rep, out, how = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        f.write(how + "\n(times under ncu are cold-cache and serialised)\n\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w:90s} {r[i]} {units[i]}\n")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
        f.write("\nwarps stalled per issue slot, by reason:\n")
        for v, h in sorted(stalls, reverse=True)[:8]:
            f.write(f"  {v:6.2f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}\n")
        hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "16"], capture_output=True, text=True).stdout
        f.write("\nhottest source lines (tools/ncu_hot.py: share of stall samples / of executed warp instructions):\n" + hot)
    print(open(out).read()[:1500])


if __name__ == "__main__":
    main()
