#!/usr/bin/env python3
"""Per-source-line totals (stall samples, instructions) from an .ncu-rep: ncu_lines.py <rep> [kernel-regex] [top N]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] == '': continue  # sass row
    try:
        line = int(r[0])
    except ValueError:
        continue
    i_s = hdr.index('# Samples'); i_i = hdr.index('Instructions Executed'); i_t = hdr.index('Thread Instructions Executed')
    k = (cur_file, line)
    a = agg.setdefault(k, [0, 0, 0, r[1]])
    a[0] += int(r[i_s] or 0); a[1] += int(r[i_i] or 0); a[2] += int(r[i_t] or 0)
tot_s = sum(a[0] for a in agg.values()) or 1; tot_i = sum(a[1] for a in agg.values()) or 1
print('total samples %d, warp instructions %d' % (tot_s, tot_i))
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print('%5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%d  %s' % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, a[2] / max(a[1], 1), f, l, a[3].strip()[:110]))
