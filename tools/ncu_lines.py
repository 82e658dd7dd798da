#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line:
samples, executed instructions, shared-memory wavefronts.  usage: ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) > 10 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[2] != '-':   # a SASS row; the line rows (Address == '-') already aggregate them
        continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur_file, int(r[0]), r[1].strip()[:90])

    def num(k):
        try:
            return float(d.get(k, 0) or 0)
        except ValueError:
            return 0.0
    agg[key] = (num('# Samples'), num('Instructions Executed'), num('L1 Wavefronts Shared'),
                num('L1 Wavefronts Shared Excessive'), num('stall_barrier'), num('stall_wait'),
                num('stall_short_sb'), num('stall_math'))
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print('total samples %d, warp instructions %d' % (tot, toti))
print('%6s %6s %6s %9s %9s  bar/wait/ssb/math  line' % ('smp%', 'inst%', 'cum%', 'smemWF', 'excess'))
cum = 0
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    cum += v[0]
    print('%6.2f %6.2f %6.1f %9d %9d  %4.1f/%4.1f/%4.1f/%4.1f  %s:%d %s' % (
        100 * v[0] / tot, 100 * v[1] / toti, 100 * cum / tot, v[2], v[3],
        100 * v[4] / tot, 100 * v[5] / tot, 100 * v[6] / tot, 100 * v[7] / tot, key[0], key[1], key[2]))
