#!/usr/bin/env python
"""One-screen digest of a bench.py JSON line (file argument or stdin)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
l = json.loads(txt.strip().splitlines()[-1])
k = l['roofline']['kernels_ms']
print('value %s %s  ms/step %s  e2e %s  launches %s  roofline frac %s' % (l['value'], l['unit'], l['ms_per_step'], l['e2e']['value'],
                                                                        l['gpu_launches'], l['roofline']['frac']))
print('kernels_ms', {n: k[n] for n in list(k)[:8]})
for c in ('config2b', 'config2c', 'config3', 'config3_cudnn_encoder', 'config4'):
    if c in l and isinstance(l[c], dict):
        print(c, {n: l[c][n] for n in ('images_per_s', 'objects_per_s', 'ms_per_step') if n in l[c]})
