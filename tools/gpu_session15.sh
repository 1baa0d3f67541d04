#!/bin/bash
mkdir -p gpurun_out
{
for cfg in "128,64,32,32 bf16 neither" "128,64,32,32 bf16 both" "512,32,32,32 f32 both" "512,64,16,16 f32 both" "512,128,8,8 f32 both" "64,256,56,56 f32 neither" "256,64,56,56 f32 both" "256,256,56,56 f32 neither" "256,3,224,224 f32 both"; do set -- $cfg
  timeout 120 python tools/perf_cabi.py crossnorm $1 $2 $3 50
  CNSN_CROSSNORM_IMPL=v1 timeout 120 python tools/perf_cabi.py crossnorm $1 $2 $3 50
done
for cfg in "128,64,32,32 bf16" "512,32,32,32 f32" "512,64,16,16 f32" "512,128,8,8 f32" "256,256,56,56 f32" "4,16,8,8 f32"; do set -- $cfg
  timeout 120 python tools/perf_cabi.py selfnorm $1 $2 neither 50
  CNSN_SELFNORM_IMPL=v1 timeout 120 python tools/perf_cabi.py selfnorm $1 $2 neither 50
done
} > gpurun_out/s15_perf.log 2>&1
cat gpurun_out/s15_perf.log
