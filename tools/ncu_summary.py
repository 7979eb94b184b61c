#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page raw --csv` output:  python tools/ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx:
            print('  %-72s %s %s' % (w, r[idx[w]][:80], units[idx[w]]))
    st = [(hdr[i], r[i]) for i in range(len(hdr)) if 'warp_issue_stalled' in hdr[i] and hdr[i].endswith('per_warp_active.pct')]
    for h, v in sorted(st, key=lambda t: -float(t[1] or 0))[:7]:
        print('     stall %-66s %s' % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
