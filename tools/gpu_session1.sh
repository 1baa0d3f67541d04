#!/bin/bash
# one GPU call: parity suite, A/B of SelfNorm implementations, default bench
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s1_pytest.log
{
for shp in "256,256,56,56 f32" "256,256,56,56 bf16" "256,512,28,28 f32" "512,32,32,32 f32"; do
  set -- $shp
  CNSN_SELFNORM_IMPL=v1 timeout 120 python tools/perf_selfnorm.py $1 $2 20
  timeout 120 python tools/perf_selfnorm.py $1 $2 20
  CNSN_SELFNORM_BWD=fused timeout 120 python tools/perf_selfnorm.py $1 $2 20
  CNSN_SELFNORM_IMPL=cluster CNSN_CLUSTER_DEBUG=1 timeout 120 python tools/perf_selfnorm.py $1 $2 20
done
} > gpurun_out/s1_ab.log 2>&1
timeout 300 python bench.py --no-train > gpurun_out/s1_bench.log 2>&1
tail -3 gpurun_out/s1_pytest.log; cat gpurun_out/s1_ab.log; tail -2 gpurun_out/s1_bench.log
