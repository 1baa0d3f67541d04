#!/bin/bash
mkdir -p gpurun_out
F="CNSN_SELFNORM_IMPL=flow"
cfgs=()
for tpi in 32 64 128 256; do for d in 3 4 6 8 12; do for keep in 0 1; do for ord in 0 1; do
  cfgs+=("$F CNSN_FLOW_TPI=$tpi CNSN_FLOW_D=$d CNSN_FLOW_KEEP=$keep CNSN_FLOW_ORDER=$ord")
done; done; done; done
timeout 600 python tools/sweep_selfnorm.py 256,256,56,56 f32 12 "-" "${cfgs[@]}" > gpurun_out/s4_sweep.log 2>&1
sort -t'|' -k3 gpurun_out/s4_sweep.log | head -3
