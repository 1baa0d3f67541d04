#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle or north_star" 2>&1 | tail -15) > gpurun_out/s2_pytest.log
tail -5 gpurun_out/s2_pytest.log
F="CNSN_SELFNORM_IMPL=flow"
timeout 300 python tools/sweep_selfnorm.py 256,256,56,56 f32 15 "-" "CNSN_SELFNORM_IMPL=v1" \
  "$F CNSN_FLOW_DEBUG=1" "$F CNSN_FLOW_D=1" "$F CNSN_FLOW_D=2" "$F CNSN_FLOW_D=4" "$F CNSN_FLOW_D=6" "$F CNSN_FLOW_D=8" "$F CNSN_FLOW_D=12" "$F CNSN_FLOW_D=16" \
  "$F CNSN_FLOW_ORDER=1" "$F CNSN_FLOW_ORDER=1 CNSN_FLOW_D=2" "$F CNSN_FLOW_ORDER=1 CNSN_FLOW_D=4" "$F CNSN_FLOW_ORDER=1 CNSN_FLOW_D=6" \
  "$F CNSN_FLOW_KEEP=1" "$F CNSN_FLOW_KEEP=1 CNSN_FLOW_D=2" "$F CNSN_FLOW_KEEP=1 CNSN_FLOW_D=6" \
  "$F CNSN_FLOW_TPI=128" "$F CNSN_FLOW_TPI=64" "$F CNSN_FLOW_TPI=32" 2>&1 | tee gpurun_out/s2_sweep.log
timeout 200 python tools/sweep_selfnorm.py 256,256,56,56 bf16 15 "CNSN_SELFNORM_IMPL=v1" "$F" "$F CNSN_FLOW_D=4" "$F CNSN_FLOW_D=12" "$F CNSN_FLOW_ORDER=1" 2>&1 | tee -a gpurun_out/s2_sweep.log
timeout 200 python tools/sweep_selfnorm.py 256,512,28,28 f32 15 "CNSN_SELFNORM_IMPL=v1" "$F" "$F CNSN_FLOW_D=8" "$F CNSN_FLOW_D=24" "$F CNSN_FLOW_ORDER=1" 2>&1 | tee -a gpurun_out/s2_sweep.log
timeout 200 python tools/sweep_selfnorm.py 256,1024,14,14 f32 15 "CNSN_SELFNORM_IMPL=v1" "$F" "$F CNSN_FLOW_D=16" "$F CNSN_FLOW_ORDER=1" 2>&1 | tee -a gpurun_out/s2_sweep.log
timeout 200 python tools/sweep_selfnorm.py 512,32,32,32 f32 15 "CNSN_SELFNORM_IMPL=v1" "-" "$F" "$F CNSN_FLOW_D=4" "$F CNSN_FLOW_ORDER=1" 2>&1 | tee -a gpurun_out/s2_sweep.log
