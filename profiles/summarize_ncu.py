#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the short text summary kept under
profiles/:  python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed_pipe_tmem.sum',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
]


def main():
  rep = sys.argv[1]
  out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  head, units = rows[0], rows[1]
  name_col = head.index('Kernel Name')
  for r in rows[2:]:
    print('kernel:', r[name_col])
    for m in METRICS:
      if m in head:
        i = head.index(m)
        print('  %-82s %s %s' % (m, r[i], units[i]))
    print()


if __name__ == '__main__':
  main()
