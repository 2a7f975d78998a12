#!/bin/bash
# Full GPU validation on a B200 box: the GPU test suite, smoke(), the full bench line.  Outputs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_tail.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py "$@" > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 400 gpurun_out/bench_full.err
python - <<'PY'
import json
l = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k: l[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, l['e2e'], l['roofline'])
oc = l.get('other_configs') or l['config'].get('other_configs')
print(json.dumps(oc)[:3500])
PY
