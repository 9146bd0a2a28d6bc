#!/bin/bash
# Race / memory checking of the hand-written kernels with compute-sanitizer on ONE GPU (run under gpurun).
# The reference has no race detection at all; this is the device-side equivalent of a TSAN pass.
set -u
mkdir -p gpurun_out
export PSB200_DEVICE_TIMEOUT=120
K="test_encode_gather_sgd and (identity or topk_f32 or scale_i8) and 3 and dtype0"
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --kernel-name kernel_substring=psb_ \
    python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$K or test_topk_wire_is_exact or test_adam_steps" \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | tail -4
done
