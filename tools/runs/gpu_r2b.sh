#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; tail -8 gpurun_out/r2b_pytest.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python tools/sweep_selfnorm.py 128,128,48,48 f32 20 "-" "i3=0" > gpurun_out/r2b_sweep.log 2>&1
timeout 200 python tools/sweep_selfnorm.py 256,128,56,56 bf16 20 "-" "i3=0" >> gpurun_out/r2b_sweep.log 2>&1
timeout 200 python tools/sweep_selfnorm.py 256,64,80,80 bf16 20 "-" "i3=0" >> gpurun_out/r2b_sweep.log 2>&1
timeout 200 python tools/sweep_selfnorm.py 512,128,40,40 f32 20 "-" "i3=0" >> gpurun_out/r2b_sweep.log 2>&1; cat gpurun_out/r2b_sweep.log
cat gpurun_out/model_parity.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --no-train 2>&1 | tail -c 600
