#!/bin/bash
# ncu launch list (cold-cache, serialised) of one GlobalGenerator forward: every k_conv_tc launch with grid and duration.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file /tmp/gen_launches.csv \
    python scripts/bench_generator.py --steps 1 > gpurun_out/gen_ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/gen_launches.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn, mn, mv, mu, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('ID')
launch = {}
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    d = launch.setdefault(int(r[idc]), {'name': r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:40]})
    v = float(r[mv].replace(',', ''))
    if r[mn] == 'gpu__time_duration.sum':
        d['us'] = v / 1e3 if r[mu] in ('ns', 'nsecond') else v
    else:
        d['grid'] = int(v)
ids = sorted(launch)
# the LAST forward: walk back from the end to the last k_reflect_pad... simply print the last 110 launches of ours
ours = [launch[i] for i in ids if launch[i]['name'].startswith('k_')]
last = ours[-75:]
tot = sum(d.get('us', 0) for d in last)
print('last %d launches of the library, %.1f us' % (len(last), tot))
for d in last:
    print('%-40s grid %6d  %8.1f us' % (d['name'], d.get('grid', 0), d.get('us', 0)))
PY
