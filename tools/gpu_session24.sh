#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_jsd.py -q -m gpu 2>&1 | tail -12) > gpurun_out/s24_pytest.log
cat gpurun_out/s24_pytest.log | tail -12
{
for cfg in "768,256,56,56 bf16" "768,512,28,28 bf16" "768,1024,14,14 bf16" "768,2048,7,7 bf16"; do set -- $cfg
  timeout 200 python tools/perf_cabi.py block $1 $2 neither 10 2>&1 | head -1
done
timeout 100 python tools/perf_cabi.py crossnorm 768,3,224,224 f32 neither 10
} > gpurun_out/s24_perf.log 2>&1
cat gpurun_out/s24_perf.log
