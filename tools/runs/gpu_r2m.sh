#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_memory or north_star" 2>&1 | tail -12
timeout 300 python tools/sweep_selfnorm.py 256,256,56,56 f32 20 "-" "tm=0" "pf=0" "pf=592" "pf=148" 2>&1 | tee gpurun_out/r2m_sweep.log
timeout 200 python tools/sweep_selfnorm.py 256,256,56,56 bf16 20 "-" "tm=0" 2>&1 | tee -a gpurun_out/r2m_sweep.log
timeout 200 python tools/sweep_selfnorm.py 512,128,40,40 f32 20 "-" "tm=0" 2>&1 | tee -a gpurun_out/r2m_sweep.log
