#!/usr/bin/env python3
"""Prints selected metrics of an .ncu-rep (ncu -i REP --page raw --csv) as `metric value unit` lines."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput',
        'dram__throughput.avg.pct', 'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'launch__occupancy',
        'sm__throughput.avg.pct', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block', 'sm__cycles_elapsed.avg ', 'sm__cycles_active.avg',
        'launch__waves_per_multiprocessor', 'sm__pipe_tensor', 'sm__inst_executed_pipe_tensor', 'smsp__cycles_active.avg ',
        'lts__t_sector_hit_rate', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'sm__pipe_alu', 'sm__pipe_fma',
        'sm__pipe_fp64', 'sm__inst_executed_pipe_xu', 'smsp__warp_issue_stalled', 'smsp__average_warps_issue_stalled',
        'sm__ctas_launched', 'smsp__warps_eligible', 'smsp__pcsamp_warps_issue_stalled', 'dram__cycles_active',
        'lts__t_bytes', 'sm__warps_active.avg.per_cycle_active', 'l1tex__data_bank_conflicts', 'smsp__inst_executed_op_shared']


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '?')[:100], 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
        for i, k in enumerate(hdr):
            if any(w.strip() in k for w in KEYS + extra) and r[i] not in ('', '0', 'n/a'):
                print(f'   {k:90s} {r[i]:>16s} {units[i]}')


if __name__ == '__main__':
    main()
