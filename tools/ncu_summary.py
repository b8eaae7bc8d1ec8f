#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --set full) and ncu launch lists into the text files kept under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  > profiles/rN_ncu_<kernel>.txt
  python tools/ncu_summary.py list gpurun_out/launches.csv  > profiles/rN_launches.txt
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {d.get('Kernel Name')}")
        for k in KEYS:
            if k in d:
                print(f"{k:100s} {d[k]:>16s} {u.get(k, '')}")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0, r[ui]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':50s} {'launches':>8s} {'total':>14s} {'avg':>12s} share")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:50s} {a[0]:8d} {a[1]:14.0f} {a[1] / a[0]:12.0f} {a[2]} {a[1] / tot * 100:5.1f}%")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
