#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_ibn.py -q -m gpu 2>&1 | tail -25) > gpurun_out/s30_pytest.log
tail -25 gpurun_out/s30_pytest.log | cut -c1-220
{
for shp in 256,64,56,56 256,128,28,28 256,256,14,14 256,256,56,56; do timeout 120 python tools/perf_ibn.py $shp 30; done
} > gpurun_out/s30_perf.log 2>&1
cat gpurun_out/s30_perf.log
