#!/usr/bin/env python
"""Opcode mix (weighted by executed count) of an ncu source-page export in SASS view.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_sass.py sass.csv [units]      # units: divide counts by this (e.g. tiles)
"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hi]
    ia = hdr.index('Instructions Executed')
    isrc = hdr.index('Source')
    ismp = hdr.index('# Samples')
    ops = defaultdict(lambda: [0.0, 0.0])
    tot = 0.0
    tots = 0.0
    for r in rows[hi + 1:]:
        try:
            n = float(r[ia])
            sm = float(r[ismp])
        except (ValueError, IndexError):
            continue
        parts = r[isrc].strip().split()
        if not parts:
            continue
        op = parts[0]
        if op.startswith('@') and len(parts) > 1:
            op = parts[1]
        op = op.split('.')[0]
        ops[op][0] += n
        ops[op][1] += sm
        tot += n
        tots += sm
    print(f'total executed {tot:.0f} ({tot / units:.1f} per unit), static instructions {len(rows) - hi - 1}')
    for op, (n, sm) in sorted(ops.items(), key=lambda x: -x[1][0])[:45]:
        print(f'{op:12s} {100 * n / tot:6.2f}%  per unit {n / units:8.1f}   samples {100 * sm / tots:5.1f}%')


if __name__ == '__main__':
    main()
