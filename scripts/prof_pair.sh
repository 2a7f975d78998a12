#!/bin/bash
# ncu --set full of the 1024-channel 3x3 convolution as a CTA pair (cta_group::2) and as 128-wide single-CTA tiles.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r2_prof_conv_pair.txt
for v in "SDN_CONV_PAIR=1" "SDN_CONV_PAIR=0 SDN_CONV_TN=256" "SDN_CONV_PAIR=0 SDN_CONV_TN=128"; do
  env $v timeout 300 ncu --set full --clock-control none -k regex:k_conv_tc -s 4 -c 1 -o /tmp/pair -f python scripts/prof_conv_big.py child > /dev/null 2>&1
  echo "## $v  (3x3, 1024 -> 1024 channels, [1,34,98,1024] -> [1,32,96,1024], no statistics)" >> gpurun_out/r2_prof_conv_pair.txt
  python scripts/ncu_summary.py /tmp/pair.ncu-rep 2>&1 | grep -E "^==|gpu__time_duration|dram__bytes_read|lts__t_bytes|launch__grid_size|launch__registers|issue_active|pipe_tensor_cycles_active_realtime.avg.pct|hmma_cycles_active_realtime|sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active|xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum |lts__t_sector_hit_rate" >> gpurun_out/r2_prof_conv_pair.txt
done
cat gpurun_out/r2_prof_conv_pair.txt
