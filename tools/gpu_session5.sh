#!/bin/bash
mkdir -p gpurun_out
F="CNSN_SELFNORM_IMPL=flow"
(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle or north_star" 2>&1 | tail -5) > gpurun_out/s5_pytest.log
tail -2 gpurun_out/s5_pytest.log
: > gpurun_out/s5_sweep.log
run() { shp=$1; dt=$2; shift 2
  cfgs=()
  for b in 1 2 3 6 12; do for la in 24 40 64; do cfgs+=("$F CNSN_FLOW_BATCHES=$b CNSN_FLOW_LOOKAHEAD_MB=$la"); done; done
  timeout 300 python tools/sweep_selfnorm.py $shp $dt 12 "CNSN_SELFNORM_IMPL=v1" "-" "${cfgs[@]}" "$F CNSN_FLOW_ORDER=1" "$F CNSN_FLOW_KEEP=0" >> gpurun_out/s5_sweep.log 2>&1
}
run 256,256,56,56 f32
run 256,256,56,56 bf16
run 256,512,28,28 f32
run 256,1024,14,14 f32
run 512,32,32,32 f32
run 512,64,16,16 f32
run 512,128,8,8 f32
run 128,64,32,32 bf16
run 64,256,56,56 f32
run 32,64,112,112 f32
