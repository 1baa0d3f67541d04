#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/perf_bn.py 2>&1 | tee gpurun_out/r2t_bn.log
timeout 300 python tools/sweep_selfnorm.py 256,2048,7,7 f32 20 "-" "grp_kb=10" "grp_kb=14" "grp_kb=28" "grp_kb=40" "poll_ns=40" 2>&1 | tee gpurun_out/r2t_grp.log
timeout 300 python tools/sweep_selfnorm.py 768,1024,14,14 bf16 20 "-" "grp_kb=10" "grp_kb=14" "grp_kb=28" "grp_kb=40" 2>&1 | tee -a gpurun_out/r2t_grp.log
timeout 300 python tools/sweep_selfnorm.py 768,2048,7,7 bf16 20 "-" "grp_kb=10" "grp_kb=40" 2>&1 | tee -a gpurun_out/r2t_grp.log
