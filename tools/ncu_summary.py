#!/usr/bin/env python
"""Prints the headline counters of one kernel from an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    for w in want:
        if w in d:
            print("%-70s %-14s %s" % (w, units[hdr.index(w)], d[w]))
    tot = sum(float(d[s].replace(",", "") or 0) for s in stalls)
    print("stall reasons (pc samples, share):")
    for s in sorted(stalls, key=lambda s: -float(d[s].replace(",", "") or 0))[:10]:
        v = float(d[s].replace(",", "") or 0)
        print("   %-32s %10.0f  %5.1f%%" % (s.replace("smsp__pcsamp_warps_issue_stalled_", ""), v, 100 * v / max(tot, 1)))
    print()
