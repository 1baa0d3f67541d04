#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cn_res" -s 6 -c 2 -o gpurun_out/r01_crossnorm_full python tools/perf_cabi.py crossnorm 256,256,56,56 f32 neither 2 > gpurun_out/s25_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sn_flow|k_sn_res" -s 6 -c 2 -o gpurun_out/r01_block_full python tools/perf_cabi.py block 256,256,56,56 f32 neither 2 > gpurun_out/s25_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_bench_launches.csv
(timeout 300 python -m pytest tests/test_jsd.py -q -m gpu 2>&1 | tail -2)
