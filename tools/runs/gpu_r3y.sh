#!/bin/bash
# round 2, step y: evidence for the channels-last kernels -- CUDA-event table, ncu launch list and details
mkdir -p gpurun_out
timeout 600 python tools/perf_nhwc.py > gpurun_out/r3y_nhwc_perf.txt 2>&1; echo "perf rc=$?"; tail -12 gpurun_out/r3y_nhwc_perf.txt
PERF_REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nhwc_|k_bn_nhwc_" -c 24 -o gpurun_out/r3y_nhwc python tools/perf_nhwc.py 512,32,32,32 f32 > gpurun_out/r3y_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r3y_nhwc.ncu-rep --page details --csv > gpurun_out/r3y_nhwc_ncu_details.csv 2>/dev/null
ls -la gpurun_out/r3y_nhwc.ncu-rep; wc -l gpurun_out/r3y_nhwc_ncu_details.csv
