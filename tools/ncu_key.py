#!/usr/bin/env python
"""Prints the handful of `ncu --page raw --csv` columns that decide what bounds a kernel (one line per metric, one value per launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
keys += [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
for k in keys:
    if k in hdr:
        i = hdr.index(k); vals = [r[i] for r in data]
        try:
            if 'stalled' in k and all(abs(float(v)) < 0.3 for v in vals): continue
        except ValueError:
            pass
        print(k.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''), units[i], [v[:60] for v in vals])
