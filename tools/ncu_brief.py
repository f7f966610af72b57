#!/usr/bin/env python3
"""usage: ncu_brief.py report.ncu-rep  -- the handful of counters that say what bounds each captured launch"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
r = list(csv.reader(io.StringIO(raw))); h = r[0]
W = [("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("inst", "smsp__inst_executed.sum"), ("occ_regs", "launch__occupancy_limit_registers"), ("occ_smem", "launch__occupancy_limit_shared_mem"),
     ("bankconf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]
ST = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle", "no_instruction", "dispatch_stall", "branch_resolving", "membar", "sleeping", "tex_throttle", "drain", "imc_miss"]
for line in r[2:]:
    print(line[h.index("Kernel Name")][:60])
    print("   " + "  ".join(f"{n}={line[h.index(m)]}" for n, m in W if m in h))
    st = []
    for s in ST:
        m = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if m in h:
            st.append((float(line[h.index(m)]), s))
    print("   stalls/issue: " + "  ".join(f"{s}={v:.2f}" for v, s in sorted(st, reverse=True)[:7]))
