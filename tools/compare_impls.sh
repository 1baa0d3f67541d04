#!/bin/bash
# A/B table: three-kernel path vs persistent kernels vs default (dataflow kernel), forward and backward, per shape.
for shp in "256,256,56,56 f32" "256,256,56,56 bf16" "256,512,28,28 f32" "256,1024,14,14 f32" "256,2048,7,7 f32" "512,32,32,32 f32" "512,64,16,16 f32" "512,128,8,8 f32" "128,64,32,32 bf16"; do
  set -- $shp
  CNSN_SELFNORM_IMPL=v1 timeout 120 python tools/perf_selfnorm.py $1 $2 20
  CNSN_SELFNORM_IMPL=persistent timeout 120 python tools/perf_selfnorm.py $1 $2 20
  timeout 120 python tools/perf_selfnorm.py $1 $2 20
done
