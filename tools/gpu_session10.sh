#!/bin/bash
mkdir -p gpurun_out
R="CNSN_FLOW_MODE=res"
timeout 200 python tools/sweep_selfnorm.py 256,256,56,56 f32 12 "-" "$R" "$R CNSN_FLOW_POLL_NS=200" "$R CNSN_FLOW_POLL_NS=1000" "$R CNSN_FLOW_POLL_NS=2000" "$R CNSN_FLOW_ORDER=1 CNSN_FLOW_POLL_NS=1000" 2>&1 | tee gpurun_out/s10_sweep.log
timeout 120 python tools/trace_flow.py 256,256,56,56 fwd gpurun_out/flow_trace_fwd.bin 2>&1 | tee gpurun_out/s10_trace_fwd.log
timeout 120 python tools/trace_flow.py 256,256,56,56 bwd gpurun_out/flow_trace_bwd.bin 2>&1 | tee gpurun_out/s10_trace_bwd.log
rm -f gpurun_out/flow_trace_fwd.bin gpurun_out/flow_trace_bwd.bin
