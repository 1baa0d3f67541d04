#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/sweep_selfnorm.py 256,2048,7,7 f32 20 "-" 2>&1 | tee gpurun_out/r3b_grp.log
timeout 200 python tools/sweep_selfnorm.py 768,1024,14,14 bf16 20 "-" 2>&1 | tee -a gpurun_out/r3b_grp.log
timeout 200 python tools/sweep_selfnorm.py 768,2048,7,7 bf16 20 "-" 2>&1 | tee -a gpurun_out/r3b_grp.log
timeout 200 python tools/sweep_selfnorm.py 256,512,7,7 f32 20 "-" 2>&1 | tee -a gpurun_out/r3b_grp.log
