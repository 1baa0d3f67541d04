#!/bin/bash
# round 2, step t: channels-last BatchNorm2d + ReLU (csrc/bn_nhwc.cu) -- parity, WideResNet-40-2 bench, per-kernel profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ibn.py -m gpu -q -k "channels_last" > gpurun_out/r3t_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3t_tests.log
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -k "wrn40_2" 2>&1 | tail -3
timeout 900 python - > gpurun_out/r3t_wrn.log 2>&1 <<'PY'
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200 import train
dev = torch.device("cuda", 0)
for cl in (True, False, True):
    r = train.bench_wrn(dev, 1, 0, steps=40, warmup=8, cn_prob=0.25, fuse_post=True, channels_last=cl)
    print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "final_loss", "memory_format", "cnsn_kernel_launches")}), flush=True)
    torch.cuda.empty_cache()
PY
echo "wrn rc=$?"; cut -c1-260 gpurun_out/r3t_wrn.log | tail -4
timeout 300 python tools/debug/wrn_profile.py benchmark cl > gpurun_out/r3t_wrnprof_cl.log 2>&1
grep "ms/step" gpurun_out/r3t_wrnprof_cl.log
