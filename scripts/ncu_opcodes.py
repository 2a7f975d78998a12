#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions) of one kernel in an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys
from collections import Counter
path, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
isrc, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = []
for r in rows[hi + 1:]:
    if len(r) > ie and r[ie].isdigit():
        data.append((r[isrc].strip(), int(r[ie]), int(r[isamp]) if r[isamp].isdigit() else 0))
    elif r and r[0] == 'Kernel Name':
        break
tot = sum(d[1] for d in data)
print('kernel', kern, 'warp instructions', tot, 'sass lines', len(data))
c, smp = Counter(), Counter()
for s, e, sa in data:
    t = s.split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += e
    smp[op] += sa
ts = sum(smp.values()) or 1
for k, v in c.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 24):
    print('  %-10s %6.2f%% of instructions   %6.2f%% of stall samples' % (k, 100.0 * v / tot, 100.0 * smp[k] / ts))
