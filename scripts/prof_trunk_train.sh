#!/bin/bash
# ncu evidence for the encoder trunk's training path (B = 16): launch list of one fwd+bwd, --set full of the weight-gradient
# kernel.  Summaries under gpurun_out/ (copied to profiles/ by hand); the .ncu-rep files stay in /tmp on the box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export ONLY=native STEPS=2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_trunk_train.csv \
    python scripts/bench_encoder_train.py > gpurun_out/trunk_ncu1.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2_launches_trunk_train.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
tot = collections.OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(',', '')); u = r[mu]
    us = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    name = r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', '')[:70]
    t = tot.setdefault(name, [0.0, 0]); t[0] += us; t[1] += 1; n += 1
half = {k: (v[0] / 2, v[1] // 2) for k, v in tot.items()}       # two steps were profiled
s = sum(v[0] for v in half.values())
print('launches per fwd+bwd: %d, serialised cold-cache device time %.1f us' % (n // 2, s))
for k, v in sorted(half.items(), key=lambda kv: -kv[1][0])[:28]:
    print('%-72s %8.1f us %5.1f%% %4d launches' % (k, v[0], 100 * v[0] / s, v[1]))
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_wgrad -s 27 -c 3 -o /tmp/wgrad -f \
    python scripts/bench_encoder_train.py > gpurun_out/trunk_ncu2.log 2>&1
python scripts/ncu_summary.py /tmp/wgrad.ncu-rep 2>&1 | grep -v "tma_red\|tma_st\|dshared\|stalled_\(membar\|sleeping\|tex\|drain\)" > gpurun_out/r2_prof_conv_wgrad.txt
head -150 gpurun_out/r2_prof_conv_wgrad.txt | grep "==\|duration\|dram__bytes\|lts__t_bytes\|tensor\|sm__throughput\|registers\|grid_size\|issue_active"
