#!/usr/bin/env python
"""Key metrics of an ncu raw-page CSV export:  ncu -i rep --page raw --csv > raw.csv"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or ('issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h):
            try:
                if 'issue_stalled' in h and float(v) < 0.05:
                    continue
            except ValueError:
                pass
            print(f'{h:92s} {u:16s} {v}')


if __name__ == '__main__':
    main()
