#!/usr/bin/env python3
"""Per-source-line totals (stall samples with their main reasons, instructions) from an .ncu-rep:
   ncu_lines.py <rep> [top N] [kernel-regex]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass']
if len(sys.argv) > 3:
    cmd += ['-k', 'regex:' + sys.argv[3], '--launch-count', '1']
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None
hdr = None
REASONS = ['stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_lg', 'stall_branch_resolving',
           'stall_not_selected', 'stall_selected', 'stall_no_inst', 'stall_membar', 'stall_sleep', 'stall_dispatch']
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] in ('File Path', 'File Name'):
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        idx = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == '':
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    a = agg.setdefault((cur_file, line), [0, 0, 0, r[1], collections.Counter()])
    try:
        a[0] += int(r[idx['# Samples']] or 0)
        a[1] += int(r[idx['Instructions Executed']] or 0)
        a[2] += int(r[idx['Thread Instructions Executed']] or 0)
        for k in REASONS:
            if k in idx:
                a[4][k] += int(r[idx[k]] or 0)
    except ValueError:
        pass  # a source line whose quotes broke the CSV row
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
allr = collections.Counter()
for a in agg.values():
    allr.update(a[4])
print('total samples %d, warp instructions %d; reasons: %s' % (tot_s, tot_i, ' '.join('%s %.1f%%' % (k[6:], 100.0 * v / tot_s) for k, v in allr.most_common(6))))
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    why = ' '.join('%s %d%%' % (k[6:], round(100.0 * v / max(a[0], 1))) for k, v in a[4].most_common(2) if v)
    print('%5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%d  [%s]  %s' % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, a[2] / max(a[1], 1), f, l, why, a[3].strip()[:90]))
