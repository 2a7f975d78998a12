#!/bin/bash
# ncu --set full of the forward rasteriser on the config-3 scenes (16 real ShapeNet templates, 768x768 samples).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
K=${1:-k_raster_tiles}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -s 2 -c 1 -o /tmp/cfg3_$K -f \
    python scripts/bench_train_step.py --native-encoder --steps 2 > gpurun_out/cfg3_ncu.log 2>&1
tail -3 gpurun_out/cfg3_ncu.log
python scripts/ncu_summary.py /tmp/cfg3_$K.ncu-rep > gpurun_out/cfg3_${K}_summary.txt 2>&1
python scripts/ncu_lines.py /tmp/cfg3_$K.ncu-rep $K 45 > gpurun_out/cfg3_${K}_lines.txt 2>&1
head -60 gpurun_out/cfg3_${K}_summary.txt
cat gpurun_out/cfg3_${K}_lines.txt
