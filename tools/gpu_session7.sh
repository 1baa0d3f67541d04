#!/bin/bash
mkdir -p gpurun_out
(CNSN_FLOW_TMA=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle or north_star or selfnorm" 2>&1 | tail -8) > gpurun_out/s7_pytest.log
tail -3 gpurun_out/s7_pytest.log
: > gpurun_out/s7_sweep.log
run() { shp=$1; dt=$2; shift 2
  cfgs=()
  for kb in 13 26 52 104; do for la in 12 20 28 40; do cfgs+=("CNSN_FLOW_TMA=1 CNSN_FLOW_ITEM_KB=$kb CNSN_FLOW_LOOKAHEAD_MB=$la"); done; done
  timeout 300 python tools/sweep_selfnorm.py $shp $dt 12 "-" "CNSN_FLOW_LOOKAHEAD_MB=32" "CNSN_FLOW_LOOKAHEAD_MB=48" "${cfgs[@]}" "CNSN_FLOW_TMA=1 CNSN_FLOW_ORDER=1" "CNSN_FLOW_TMA=1 CNSN_FLOW_DEBUG=1" >> gpurun_out/s7_sweep.log 2>&1
}
run 256,256,56,56 f32
run 256,256,56,56 bf16
run 256,512,28,28 f32
run 256,1024,14,14 f32
run 512,32,32,32 f32
