#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "channels_last" > gpurun_out/r3q_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3q_tests.log
timeout 600 python tools/debug/wrn_host_profile.py > gpurun_out/r3q_host.log 2>&1; echo "host rc=$?"
grep "ms/step" gpurun_out/r3q_host.log
