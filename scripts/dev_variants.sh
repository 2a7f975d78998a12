#!/bin/bash
# times bench.py's per-kernel table for each k_line_sweep variant
for v in 0 1 2; do
  SDN_LINE_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms']
print('variant $v', d['value'], d['ms_per_step'], 'line', k.get('k_line_sweep'), 'face', k.get('k_bwd_face'), 'tiles', k.get('k_raster_tiles'))"
done
