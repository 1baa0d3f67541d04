#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_site.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/perf_site.py 20 4 2>&1 | tee gpurun_out/r3g_site.log
CNSN_TUNE_TM=0 timeout 300 python tools/perf_site.py 20 4 2>&1 | tee -a gpurun_out/r3g_site.log
