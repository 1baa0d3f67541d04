#!/bin/bash
mkdir -p gpurun_out
{
for cfg in "256,256,56,56 f32" "256,512,28,28 f32" "256,1024,14,14 f32" "256,2048,7,7 f32" "512,32,32,32 f32" "256,256,56,56 bf16"; do set -- $cfg
  timeout 200 python tools/perf_cabi.py block $1 $2 neither 20 2>&1
done
} > gpurun_out/s21_block.log 2>&1
cat gpurun_out/s21_block.log
timeout 300 python tools/train_bench.py wrn 512 20 > gpurun_out/s21_wrn.log 2>&1; cat gpurun_out/s21_wrn.log | tail -4
timeout 600 python tools/train_bench.py resnet50 256 6 > gpurun_out/s21_r50.log 2>&1; cat gpurun_out/s21_r50.log | tail -6
