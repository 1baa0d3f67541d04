#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "crossnorm" 2>&1 | tail -15) > gpurun_out/s14_pytest.log
tail -6 gpurun_out/s14_pytest.log
{
for cfg in "128,64,32,32 bf16 neither" "128,64,32,32 bf16 both" "512,32,32,32 f32 both" "512,64,16,16 f32 both" "512,128,8,8 f32 both" "64,256,56,56 f32 neither" "256,64,56,56 f32 both"; do set -- $cfg
  CNSN_FLOW_DEBUG=1 timeout 120 python tools/perf_crossnorm.py $1 $2 $3 30 2>&1 | grep -v "^\[cnsn flow" 
  CNSN_FLOW_DEBUG=1 timeout 120 python tools/perf_crossnorm.py $1 $2 $3 1 2>&1 | grep "^\[cnsn flow" | sort | uniq -c
  CNSN_CROSSNORM_IMPL=v1 timeout 120 python tools/perf_crossnorm.py $1 $2 $3 30 2>&1 | head -1 | sed 's/^/   v1: /'
done
} > gpurun_out/s14_perf.log 2>&1
cat gpurun_out/s14_perf.log
