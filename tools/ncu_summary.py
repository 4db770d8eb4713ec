#!/usr/bin/env python3
"""Prints the metrics we track from an .ncu-rep (run here, no GPU needed): tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.sum.per_cycle_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum']
for i, h in enumerate(hdr):
    if h in want:
        print('%-75s %-12s %s' % (h, units[i], [r[i] for r in data]))
print('--- warp stall reasons (per issue active)')
st = []
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
        st.append((float(data[0][i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
for v, n in sorted(st, reverse=True)[:10]:
    print('   %-28s %.3f' % (n, v))
