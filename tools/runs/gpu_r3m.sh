#!/bin/bash
# round 2, step m: smoke with the BatchNorm cases, fresh step profiles (WideResNet-40-2, ResNet-50 fp32)
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r3m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3m_smoke.log
timeout 600 python -m pytest tests/test_ibn.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/debug/wrn_profile.py benchmark > gpurun_out/r3m_wrnprof.log 2>&1; echo "wrn rc=$?"
timeout 600 python tools/debug/r50_profile.py > gpurun_out/r3m_r50prof.log 2>&1; echo "r50 rc=$?"
grep "ms/step" gpurun_out/r3m_wrnprof.log gpurun_out/r3m_r50prof.log
