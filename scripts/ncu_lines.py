#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions and stall samples of one kernel (ncu --import-source on, -lineinfo)."""
import csv, subprocess, sys
path, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data, hdr = {}, None
for r in rows:
    if 'Instructions Executed' in r:
        hdr = r
        ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= ie or not r[0].isdigit() or not r[ie].isdigit():
        continue
    ln = int(r[0])
    e, sa = int(r[ie]), int(r[isamp]) if r[isamp].isdigit() else 0
    d = data.setdefault(ln, [r[1].strip(), 0, 0])
    d[1] += e; d[2] += sa
tot = sum(d[1] for d in data.values()); ts = sum(d[2] for d in data.values()) or 1
print('kernel', kern, 'warp instructions', tot)
for ln, (src, e, sa) in sorted(data.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5d %6.2f%% %6.2f%%  %s' % (ln, 100.0 * e / tot, 100.0 * sa / ts, src[:120]))
