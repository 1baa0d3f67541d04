#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_robustness_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/debug/r50_profile.py 2>&1 | cut -c1-210 | tee gpurun_out/r2w_r50prof.log | grep -v "^---" | head -36
