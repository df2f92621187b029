#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters we track.  usage: ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
STALL = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'
STALLS = ['barrier', 'long_scoreboard', 'short_scoreboard', 'mio_throttle', 'lg_throttle', 'math_pipe_throttle', 'wait',
          'not_selected', 'membar', 'dispatch_stall', 'branch_resolving', 'no_instruction', 'drain', 'tex_throttle', 'sleeping',
          'imc_miss']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('=' * 100)
        print(r[hdr.index('Kernel Name')][:110])
        for w in WANT:
            if w in hdr:
                print(f'  {w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
        st = []
        for s in STALLS:
            k = STALL % s
            if k in hdr:
                st.append((float(r[hdr.index(k)] or 0), s))
        print('  stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in sorted(st, reverse=True)[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
