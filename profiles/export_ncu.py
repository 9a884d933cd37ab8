#!/usr/bin/env python
"""Turns the `.ncu-rep` files of a profiling call (gpurun_out/, scratch) into the small CSV / text summaries
that are committed under profiles/:   python profiles/export_ncu.py <tag> <report.ncu-rep> [kernel-substring]
 -> profiles/<tag>_full.csv   selected `--page raw` metrics, one row per captured launch
 -> profiles/<tag>_lines.txt  instructions / stall samples folded onto source lines (tools/ncu_by_line.py)"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep = sys.argv[1], sys.argv[2]
pat = sys.argv[3] if len(sys.argv) > 3 else "step"
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h == "Kernel Name" or h.startswith((
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fp64.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
    "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct", "sm__pipe_tensor_cycles_active.avg.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed")) or
    (h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"))]
keep = [i for i in keep if not hdr[i].endswith((".per_second", ".pct_of_peak_sustained_elapsed")) or "throughput" in hdr[i]]
with open(os.path.join(ROOT, "profiles", tag + "_full.csv"), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep])
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), rep, pat], capture_output=True, text=True,
                     env=dict(os.environ, TOP="60")).stdout
open(os.path.join(ROOT, "profiles", tag + "_lines.txt"), "w").write(out)
print("wrote", tag + "_full.csv,", tag + "_lines.txt", f"({len(rows) - 2} launches)")
