#!/usr/bin/env python
"""One line per captured launch from the text that scripts/ncu_summary.py printed: duration, tensor-pipe %, issue %, DRAM."""
import re, sys
KEYS = [('gpu__time_duration.sum', 'us'), ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_%'),
        ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_mem_%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_%'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'), ('lts__t_sector_hit_rate.pct', 'l2_hit_%'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('smsp__inst_executed.sum', 'warp_inst'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu_%'), ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma_%')]
cur, rows = None, []
for line in open(sys.argv[1]):
    if line.startswith('== '):
        cur = {'kernel': re.sub(r'\(.*', '', line.split(' ', 2)[2]).replace('void ', '').replace('<unnamed>::', '').strip()}
        rows.append(cur)
        continue
    t = line.split()
    if cur is not None and len(t) >= 2:
        for k, short in KEYS:
            if t[0] == k:
                cur[short] = ' '.join(t[1:3]) if short.startswith('dram') else t[1]
labels = sys.argv[2:]        # optional names for the launches, in order
cols = ['kernel'] + [s for _, s in KEYS]
print('| # | ' + ' | '.join(cols) + ' |')
print('|' + '---|' * (len(cols) + 1))
for i, r in enumerate(rows):
    name = labels[i] if i < len(labels) else str(i)
    def f(v):
        try:
            return '%.1f' % float(v) if '.' in v else v
        except Exception:
            return v
    print('| ' + name + ' | ' + ' | '.join(f(r.get(c, '')) for c in cols) + ' |')
