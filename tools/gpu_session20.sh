#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "block or fused_vs_oracle or cfg1 or golden" 2>&1 | tail -12) > gpurun_out/s20_pytest.log
tail -5 gpurun_out/s20_pytest.log
{
for cfg in "256,256,56,56 f32" "256,512,28,28 f32" "256,1024,14,14 f32" "256,2048,7,7 f32" "512,32,32,32 f32"; do set -- $cfg
  CNSN_FLOW_DEBUG=1 timeout 200 python tools/perf_cabi.py block $1 $2 neither 20 2>&1 | sort | uniq -c | sort -k1,1n | sed 's/^ *[0-9]* //' | tail -4
done
} > gpurun_out/s20_perf.log 2>&1
cat gpurun_out/s20_perf.log
