#!/usr/bin/env python3
"""Summarise ncu outputs into small text files for profiles/ (read here, no GPU needed).
   ncu_summary.py raw <file.ncu-rep> <out.txt>        key counters per captured kernel
   ncu_summary.py launches <launches.csv> <out.txt>   per-kernel totals and shares of a --metrics gpu__time_duration.sum pass"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__maximum_warps_per_active_cycle_pct', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def raw(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, 'w') as f:
        f.write('# %s (ncu --set full --clock-control none; replayed, cold-cache: not a bench value)\n' % rep)
        for r in rows[2:]:
            f.write('\n== %s  grid %s block %s\n' % (r[idx['Kernel Name']], r[idx['launch__grid_size']], r[idx['launch__block_size']]))
            for w in WANT:
                if w in idx:
                    f.write('  %-88s %s %s\n' % (w, r[idx[w]], units[idx[w]]))


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
        a = agg.setdefault(row['Kernel Name'].split('(')[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write('# %s: ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold cache: compare SHARES)\n' % path)
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('%-70s launches %5d  total %12.1f us  avg %10.1f us  share %.4f\n' % (k, a[0], a[1], a[1] / a[0], a[1] / tot))


if __name__ == '__main__':
    {'raw': raw, 'launches': launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
