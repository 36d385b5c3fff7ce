#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
stall = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
for r in rows[2:]:
    print("==", r[idx['Kernel Name']][:90])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")
    vals = sorted([(float((r[idx[h]] or '0').replace(',', '')), h) for h in stall], reverse=True)
    tot = sum(v for v, _ in vals) or 1
    print("  stalls: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')}={100 * v / tot:.0f}%" for v, h in vals[:7]))
