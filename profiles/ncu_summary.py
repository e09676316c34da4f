#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: python scripts_ncu_summary.py raw.csv [substr ...]"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
hdr, units, rows = r[0], r[1], r[2:]
keys = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
    'sm__throughput.avg.pct', 'warps_active.avg.pct', 'registers_per_thread', 'occupancy_limit', 'issue_active.avg.pct',
    'issue_stalled', 'thread_inst_executed_per_inst', 'bank_conflicts', 'shared_atom', 'waves_per', 'pipe_tensor', 'inst_executed.sum',
    'pipe_lsu', 'pipe_alu', 'pipe_fma', 'pipe_xu', 'l1tex__t_bytes', 'lts__t_bytes', 'sm__cycles_active.avg', 'achieved_occupancy', 'shared_mem']
for ri, row in enumerate(rows):
    print('=== launch', ri, row[4][:80], 'grid', row[8], 'block', row[7])
    for h, u, v in zip(hdr, units, row):
        if any(k in h for k in keys) and v not in ('', '0', 'n/a'):
            print(f"  {h:100s} {v:>18s} {u}")
    break
