#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "channels_last" > gpurun_out/r3p_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3p_tests.log
timeout 300 python tools/debug/wrn_profile.py benchmark cl aug > gpurun_out/r3p_wrnprof_cl_aug.log 2>&1
timeout 300 python tools/debug/wrn_profile.py benchmark nchw aug > gpurun_out/r3p_wrnprof_nchw_aug.log 2>&1
grep "ms/step" gpurun_out/r3p_wrnprof_*.log
