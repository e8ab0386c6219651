#!/usr/bin/env python3
"""usage: scripts/ncu_summary.py <report.ncu-rep> [more metric substrings]  -> key raw metrics per launch"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'issue_stalled']
print(f'# ncu summary of {rep} (ncu --set full --clock-control none)')
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    for k in hdr:
        if any(w in k for w in want + extra):
            if 'issue_stalled' in k and not k.endswith('per_issue_active.ratio'):
                continue
            print(f'{k:100s} {d[k]:>22s} {u.get(k, "")}')
    print()
