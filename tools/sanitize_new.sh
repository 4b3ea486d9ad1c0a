#!/bin/bash
# compute-sanitizer over the kernels added in session 3 (k_msm_direct, k_precompute_direct, k_point_sum_batch, sharded accumulator)
SEL='direct and pallas and ((small and (127 or 128)) or row_sharded or empty_identity)'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool  pytest -k \"$SEL\""
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error|Error" | head -8
done
