#!/bin/bash
# compute-sanitizer memcheck over the GPU tests of the kernels added late in round 2 (conv pair form, wgrad, BatchNorm,
# max-pool, transposed-operand pack, im2col rows) + racecheck over the shared-memory-heavy ones.  Summaries -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T="tests/test_gpu_conv_grad.py tests/test_gpu_conv_tc.py tests/test_gpu_encoder_tc.py::test_training_trunk_forward_backward_matches_torchvision"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 20 python -m pytest $T -q -x -p no:cacheprovider > /tmp/memcheck.log 2>&1
{ echo "## memcheck: $T"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" /tmp/memcheck.log | head -20; } > gpurun_out/r2_sanitizer_late.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 --print-limit 20 python -m pytest "tests/test_gpu_conv_grad.py::test_max_pool_forward_and_backward_match_torch_exactly" "tests/test_gpu_conv_grad.py::test_transposed_weight_operands" "tests/test_gpu_conv_grad.py::test_batch_norm_training_forward_and_backward" -q -x -p no:cacheprovider > /tmp/racecheck.log 2>&1
{ echo "## racecheck: max-pool, transposed operands, BatchNorm kernels"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" /tmp/racecheck.log | head -12; } >> gpurun_out/r2_sanitizer_late.txt
cat gpurun_out/r2_sanitizer_late.txt
