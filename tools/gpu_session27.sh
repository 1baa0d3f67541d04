#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle or block" 2>&1 | tail -8) > gpurun_out/s27_pytest.log
tail -4 gpurun_out/s27_pytest.log
{
for cfg in "256,2048,7,7 f32" "256,1024,14,14 bf16" "256,2048,7,7 bf16" "768,1024,14,14 bf16" "768,2048,7,7 bf16"; do set -- $cfg
  CNSN_FLOW_DEBUG=1 timeout 200 python tools/perf_cabi.py selfnorm $1 $2 neither 20 2>&1 | grep -v "^\[cnsn" 
  CNSN_FLOW_DEBUG=1 timeout 200 python tools/perf_cabi.py selfnorm $1 $2 neither 1 2>&1 | grep "^\[cnsn" | sort -u
  timeout 200 python tools/perf_cabi.py block $1 $2 neither 20 2>&1 | head -1
done
} > gpurun_out/s27_perf.log 2>&1
cat gpurun_out/s27_perf.log
