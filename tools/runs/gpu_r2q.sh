#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; tail -5 gpurun_out/r2q_pytest.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-train --no-crossnorm --no-sustained > gpurun_out/r2q_ncu_list.log 2>&1; tail -2 gpurun_out/r2q_ncu_list.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sn_tm -s 4 -c 2 -f -o gpurun_out/r02_tm_full python tools/sweep_selfnorm.py 256,256,56,56 f32 1 "-" > gpurun_out/r2q_ncu_full.log 2>&1; tail -3 gpurun_out/r2q_ncu_full.log
ncu -i gpurun_out/r02_tm_full.ncu-rep --page details --csv > gpurun_out/r02_tm_ncu_details.csv 2>/dev/null; ls -la gpurun_out/r02_*
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.log 2> gpurun_out/r2q_bench.err; head -c 1200 gpurun_out/r2q_bench.log
