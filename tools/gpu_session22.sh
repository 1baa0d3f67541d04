#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle or block" 2>&1 | tail -8) > gpurun_out/s22_pytest.log
tail -3 gpurun_out/s22_pytest.log
{
for m in l2 res dyg; do
  CNSN_FLOW_BWD=$m timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 f32 neither 30
done
for pf in 0 300 1184 2400; do CNSN_FLOW_PF=$pf CNSN_FLOW_BWD=dyg timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 f32 neither 30; done
for kb in 13 50; do CNSN_FLOW_ITEM_KB=$kb CNSN_FLOW_BWD=dyg timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 f32 neither 30; done
for m in l2 res dyg; do
  CNSN_FLOW_BWD=$m timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 bf16 neither 30
  CNSN_FLOW_BWD=$m timeout 120 python tools/perf_cabi.py selfnorm 256,512,28,28 f32 neither 30
done
CNSN_FLOW_BWD=dyg timeout 120 python tools/perf_cabi.py block 256,256,56,56 f32 neither 20
} > gpurun_out/s22_perf.log 2>&1
cat gpurun_out/s22_perf.log
