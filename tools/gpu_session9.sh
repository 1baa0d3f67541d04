#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle" 2>&1 | tail -8) > gpurun_out/s9_pytest.log
tail -3 gpurun_out/s9_pytest.log
: > gpurun_out/s9_sweep.log
R="CNSN_FLOW_MODE=res"
run() { shp=$1; dt=$2; shift 2
  timeout 300 python tools/sweep_selfnorm.py $shp $dt 12 "-" "$R CNSN_FLOW_DEBUG=1" "$R CNSN_FLOW_ITEM_KB=13" "$R CNSN_FLOW_ITEM_KB=50" "$R CNSN_FLOW_ITEM_KB=100" "$R CNSN_FLOW_ORDER=1" "$R CNSN_FLOW_ORDER=1 CNSN_FLOW_ITEM_KB=50" >> gpurun_out/s9_sweep.log 2>&1
}
run 256,256,56,56 f32
run 256,256,56,56 bf16
run 256,512,28,28 f32
run 256,1024,14,14 f32
run 512,32,32,32 f32
run 512,64,16,16 f32
run 128,64,32,32 bf16
grep -v "cnsn flow" gpurun_out/s9_sweep.log; grep "cnsn flow" gpurun_out/s9_sweep.log | sort | uniq -c
