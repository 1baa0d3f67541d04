#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_binding_gpu.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/perf_host.py 300 2>&1 | tee gpurun_out/r2f_host.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
timeout 200 python tools/sweep_selfnorm.py 256,64,80,80 bf16 20 "-" "i3=0" 2>&1 | tail -2
