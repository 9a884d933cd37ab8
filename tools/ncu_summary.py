#!/usr/bin/env python
"""Development tool: the handful of `ncu --page raw` numbers that decide what a kernel is bound by.
    python tools/ncu_summary.py <report.ncu-rep> [more reports]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print("==", rep, "::", d.get("Kernel Name", "")[:90])
        for k in KEYS:
            if k in d:
                print(f"  {k:72s} {d[k]:>16s} {u[k]}")
        st = {h: float(d[h]) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and d[h]}
        tot = sum(st.values()) or 1
        print("  stalls: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%"
                                        for h, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
