#!/bin/bash
mkdir -p gpurun_out
R="CNSN_FLOW_MODE=res"
(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle" 2>&1 | tail -8) > gpurun_out/s11_pytest.log
tail -3 gpurun_out/s11_pytest.log
timeout 200 python tools/sweep_selfnorm.py 256,256,56,56 f32 12 "-" "$R" "$R CNSN_FLOW_POLL_NS=100" "$R CNSN_FLOW_POLL_NS=1000" "$R CNSN_FLOW_ITEM_KB=13" "$R CNSN_FLOW_ITEM_KB=50" "$R CNSN_FLOW_ORDER=1" 2>&1 | tee gpurun_out/s11_sweep.log
timeout 120 python tools/trace_flow.py 256,256,56,56 fwd gpurun_out/flow_trace_fwd.bin 2>&1 | tee gpurun_out/s11_trace_fwd.log
timeout 120 python tools/trace_flow.py 256,256,56,56 bwd gpurun_out/flow_trace_bwd.bin 2>&1 | tee gpurun_out/s11_trace_bwd.log
rm -f gpurun_out/flow_trace_fwd.bin gpurun_out/flow_trace_bwd.bin
for shp in "256,256,56,56 bf16" "256,512,28,28 f32" "256,1024,14,14 f32" "512,32,32,32 f32" "128,64,32,32 bf16"; do set -- $shp
timeout 200 python tools/sweep_selfnorm.py $1 $2 12 "-" "$R" "$R CNSN_FLOW_ITEM_KB=50" 2>&1 | tee -a gpurun_out/s11_sweep.log
done
