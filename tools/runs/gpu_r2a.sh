#!/bin/bash
# Round 2, first GPU call: parity of the persistent cooperative kernels, then A/B timings at the north-star shape.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; tail -6 gpurun_out/r2a_pytest.log
timeout 90 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python tools/sweep_selfnorm.py 256,256,56,56 f32 20 "-" "cooperative=0" "i3=1" "flow_bwd=res" "flow_bwd=dyg" "flow_mode=l2" "pf=0" "item_kb=38" > gpurun_out/r2a_sweep.log 2>&1; cat gpurun_out/r2a_sweep.log
timeout 200 python tools/sweep_selfnorm.py 256,2048,7,7 f32 20 "-" "cooperative=0" "selfnorm_impl=v1" > gpurun_out/r2a_sweep_grp.log 2>&1; cat gpurun_out/r2a_sweep_grp.log
timeout 200 python tools/sweep_selfnorm.py 768,1024,14,14 bf16 20 "-" "selfnorm_impl=v1" >> gpurun_out/r2a_sweep_grp.log 2>&1; tail -2 gpurun_out/r2a_sweep_grp.log
timeout 200 python tools/perf_site.py 20 0,1,2,4,5 > gpurun_out/r2a_site.log 2>&1; cat gpurun_out/r2a_site.log
timeout 200 python tools/perf_crossnorm.py > gpurun_out/r2a_cn.log 2>&1; tail -12 gpurun_out/r2a_cn.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; tail -c 2500 gpurun_out/r2a_bench.log; tail -5 gpurun_out/r2a_bench.err
