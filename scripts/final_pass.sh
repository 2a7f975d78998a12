#!/bin/bash
# round-end measurement pass on the GPU box (run under gpurun): tests, smoke, both bench arms, ncu launch lists
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -2 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_arm.json; cut -c1-400 gpurun_out/bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_object|k_linear_tc' -c 40 --csv --log-file gpurun_out/launches_train.csv python scripts/bench_train_step.py --steps 2 > gpurun_out/train_under_ncu.txt 2>&1
timeout 200 python scripts/bench_train_step.py --graph 2>/dev/null | tail -1 > gpurun_out/train_step_graph.json; cut -c1-300 gpurun_out/train_step_graph.json
wc -l gpurun_out/launches_bench.csv gpurun_out/launches_train.csv
