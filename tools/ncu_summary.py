#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one block per captured launch with the metrics the roofline uses."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'smsp__cycles_elapsed.avg.per_second', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_bytes.sum', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
if pat == '--list':
    for h, u in zip(hdr, units):
        print(h, u)
    sys.exit(0)
for r in rows[2:]:
    print('----')
    for w in (WANT if not pat else [h for h in hdr if pat in h]):
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
