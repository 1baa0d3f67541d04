#!/bin/bash
# round 2, step 4b: compute-sanitizer over tools/sanitize.py (incl. the channels-last kernels), non-finite test
mkdir -p gpurun_out
timeout 600 python tools/sanitize.py > gpurun_out/r4b_plain.log 2>&1; echo "plain rc=$?"; tail -8 gpurun_out/r4b_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/r4b_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/r4b_sanitizer_$tool.log | tail -2
done
timeout 600 python -m pytest tests/test_robustness_gpu.py -m gpu -q -k "non_finite" 2>&1 | tail -2
