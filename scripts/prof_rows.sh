#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python scripts/prof_rows.py
timeout 300 ncu --set full --clock-control none -k regex:k_conv_rows -s 3 -c 1 -o /tmp/rows -f python scripts/prof_rows.py > gpurun_out/rows_ncu.log 2>&1
ncu -i /tmp/rows.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=['gpu__time_duration.sum','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu.sum','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','sm__cycles_elapsed.avg','smsp__cycles_active.avg','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum']
for i,n in enumerate(h):
    if any(n==w or n.endswith(w) for w in want): print(n, '=', v[i])
"
