#!/usr/bin/env python
"""Aggregate the source page of an ncu report by source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_hot.py src.csv [top_n]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    cur = None
    hdr = None
    agg = defaultdict(lambda: defaultdict(float))
    text = {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or cur is None:
            continue
        d = dict(zip(hdr, r))
        try:
            line = int(r[0])
        except ValueError:
            continue
        key = (cur, line)
        text[key] = r[1].strip()[:90]
        for k in ('# Samples', 'Instructions Executed', 'stall_wait', 'stall_math', 'stall_short_sb',
                  'stall_long_sb', 'stall_selected', 'stall_not_selected', 'stall_branch_resolving',
                  'stall_no_inst', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal'):
            v = d.get(k, '')
            try:
                agg[key][k] += float(v)
            except ValueError:
                pass
    tot = sum(a['# Samples'] for a in agg.values()) or 1.0
    toti = sum(a['Instructions Executed'] for a in agg.values()) or 1.0
    print(f'total samples {tot:.0f}, instructions {toti:.0f}')
    byfile = defaultdict(float)
    for (f, l), a in agg.items():
        byfile[f] += a['# Samples']
    for f, s in sorted(byfile.items(), key=lambda x: -x[1]):
        print(f'  {f:20s} {100*s/tot:5.1f}%')
    print(f"{'file:line':22s} {'smp%':>6s} {'inst%':>6s} {'wait':>6s} {'math':>6s} {'ssb':>6s} {'lsb':>6s} {'sel':>6s} {'nsel':>6s} {'br':>5s} {'noi':>5s} {'wf/id':>6s}  source")
    for key, a in sorted(agg.items(), key=lambda x: -x[1]['# Samples'])[:top]:
        s = a['# Samples'] or 1.0
        wf = a['L1 Wavefronts Shared'] / a['L1 Wavefronts Shared Ideal'] if a['L1 Wavefronts Shared Ideal'] else 0
        print(f"{key[0][:16]+':'+str(key[1]):22s} {100*a['# Samples']/tot:6.2f} {100*a['Instructions Executed']/toti:6.2f} "
              f"{100*a['stall_wait']/s:6.1f} {100*a['stall_math']/s:6.1f} {100*a['stall_short_sb']/s:6.1f} {100*a['stall_long_sb']/s:6.1f} "
              f"{100*a['stall_selected']/s:6.1f} {100*a['stall_not_selected']/s:6.1f} {100*a['stall_branch_resolving']/s:5.1f} {100*a['stall_no_inst']/s:5.1f} {wf:6.2f}  {text[key]}")


if __name__ == '__main__':
    main()
