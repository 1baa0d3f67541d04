#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/micro/order_bw 256 256 3136 > gpurun_out/s8_order.log 2>&1
cat gpurun_out/s8_order.log
timeout 200 python tools/sweep_selfnorm.py 256,256,56,56 f32 12 "-" "CNSN_FLOW_KU=8" "CNSN_FLOW_KU=8 CNSN_FLOW_TPI=128" "CNSN_FLOW_KU=8 CNSN_FLOW_TPI=32" "CNSN_FLOW_KU=8 CNSN_FLOW_LOOKAHEAD_MB=32" "CNSN_FLOW_KU=8 CNSN_FLOW_LOOKAHEAD_MB=48"  2>&1 | tee gpurun_out/s8_sweep.log
