#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s13_pytest.log
tail -3 gpurun_out/s13_pytest.log
timeout 300 python bench.py --no-train > gpurun_out/s13_bench.log 2>&1
tail -1 gpurun_out/s13_bench.log | cut -c1-1200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_flow_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sn_res|k_sn_flow" -s 4 -c 2 -o gpurun_out/r01_flow_full python tools/sweep_selfnorm.py 256,256,56,56 f32 1 "-" > gpurun_out/s13_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
