#!/bin/bash
mkdir -p gpurun_out
{
for pf in 0 600 1184 1800 2400 4000; do
  CNSN_FLOW_PF=$pf CNSN_FLOW_MODE=res timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 f32 neither 30
done
for pf in 0 1184 2400; do
  CNSN_FLOW_PF=$pf timeout 120 python tools/perf_cabi.py crossnorm 256,256,56,56 f32 neither 30
  CNSN_FLOW_PF=$pf CNSN_FLOW_MODE=res timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 bf16 neither 30
  CNSN_FLOW_PF=$pf CNSN_FLOW_MODE=res timeout 120 python tools/perf_cabi.py selfnorm 256,512,28,28 f32 neither 30
  CNSN_FLOW_PF=$pf CNSN_FLOW_MODE=res timeout 120 python tools/perf_cabi.py selfnorm 512,32,32,32 f32 neither 30
done
} > gpurun_out/s17_perf.log 2>&1
cat gpurun_out/s17_perf.log
