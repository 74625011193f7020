#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python profiles/ncu_summary.py <rep> [kernel-regex]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
for row in rows[2:]:
    if len(sys.argv) > 2 and sys.argv[2] not in row[hdr.index("Kernel Name")]:
        continue
    print("=" * 100)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("  %-72s %s %s" % (k, row[i][:90], units[i]))
    st = sorted(((float(row[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:8]
    for v, h in st:
        print("  stall %-60s %6.1f %%" % (h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))
