#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) of one kernel into the text summary kept under profiles/:
the Speed-of-Light / occupancy / scheduler lines, DRAM bytes, and the warp-stall mix.  Usage:
    python tools/ncu_summary.py gpurun_out/mas_final.ncu-rep "title line" > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEEP = ["Duration", "Memory Throughput", "DRAM Throughput", "Compute (SM) Throughput", "Executed Ipc Active",
        "Issue Slots Busy", "L1/TEX Hit Rate", "L2 Hit Rate", "No Eligible", "Eligible Warps Per Scheduler",
        "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp", "Avg. Not Predicated Off Threads Per Warp",
        "Grid Size", "Block Size", "Registers Per Thread", "Dynamic Shared Memory Per Block", "Theoretical Occupancy",
        "Achieved Occupancy", "Branch Efficiency"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    print("# " + title)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.DictReader(io.StringIO(det)))
    if rows:
        print("  " + rows[0]["Kernel Name"] + "  grid " + rows[0].get("Grid Size", "") + " block " + rows[0].get("Block Size", ""))
    seen = set()
    for r in rows:
        n = r["Metric Name"]
        if n in KEEP and (n, r["Section Name"]) not in seen:
            seen.add((n, r["Section Name"]))
            print("    %-45s %-16s %s" % (n, r["Metric Unit"], r["Metric Value"]))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rr[0], rr[1], rr[2]
    col = {h: i for i, h in enumerate(hdr)}
    for m in RAW:
        if m in col:
            print("%s %s %s" % (m, units[col[m]], vals[col[m]]))
    stalls = {}
    for h, i in col.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(vals[i].replace(",", ""))
            except ValueError:
                pass
    tot = sum(stalls.values())
    if tot:
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:9]
        print("# stall samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in top))


if __name__ == "__main__":
    main()
