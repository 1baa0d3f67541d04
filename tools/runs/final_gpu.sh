#!/bin/bash
# Round-end style validation on one B200: full GPU parity suite, smoke, bench (both arms), launch list and ncu
# capture of the fused-site kernels.  Everything is bounded by its own timeout.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/s40_pytest.log 2>&1; tail -4 gpurun_out/s40_pytest.log
timeout 90 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/s40_bench.log 2> gpurun_out/s40_bench.err; tail -c 3000 gpurun_out/s40_bench.log
timeout 200 python bench.py --impl reference --no-train > gpurun_out/s40_ref.log 2>&1; tail -c 600 gpurun_out/s40_ref.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_site_launches.csv \
    python tools/perf_site.py 3 0,4 > gpurun_out/s40_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_site_res -s 4 -c 2 -f -o gpurun_out/r01_site_full \
    python tools/perf_site.py 3 4 > gpurun_out/s40_ncu_full.log 2>&1
ls -la gpurun_out/r01_site*
