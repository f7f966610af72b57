#!/usr/bin/env python3
"""Condense `ncu --set full` reports into the per-kernel summary committed under profiles/.
usage: ncu_summary.py out.csv report1.ncu-rep [report2.ncu-rep ...]   (needs `ncu` on PATH; reads the raw page)"""
import csv
import io
import subprocess
import sys

COLS = ["launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum"]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    rows, units = [], None
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        r = list(csv.reader(io.StringIO(raw)))
        hdr, un = r[0], r[1]
        ix = [hdr.index(c) for c in COLS]
        kn = hdr.index("Kernel Name")
        units = [un[i] for i in ix]
        for line in r[2:]:
            if len(line) > max(ix):
                rows.append([line[kn].split("(")[0].replace("gvs::", "")] + [line[i] for i in ix])
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + COLS)
        w.writerow([""] + units)
        w.writerows(rows)
    for r in rows:
        print(f"{r[0][:40]:40s} {float(r[2]):9.1f} us  dram {float(r[3]):7.2f}+{float(r[4]):6.2f} MB  alu {float(r[7]):5.1f}%  issue {float(r[11]):5.1f}%  warps {float(r[12]):5.1f}%")


if __name__ == "__main__":
    main()
