#!/bin/bash
# round 2, step o: channels-last SelfNorm (csrc/selfnorm_nhwc.cu) -- parity, WideResNet-40-2 step NCHW against channels_last
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "channels_last" > gpurun_out/r3o_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3o_tests.log
timeout 900 python - > gpurun_out/r3o_wrn.log 2>&1 <<'PY'
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200 import train
dev = torch.device("cuda", 0)
for cl in (False, True):
    for cn_prob in (0.0, 0.25):
        r = train.bench_wrn(dev, 1, 0, steps=40, warmup=8, cn_prob=cn_prob, fuse_post=True, channels_last=cl)
        print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "final_loss", "memory_format", "graph", "cnsn_kernel_launches")} | {"cn_prob": cn_prob}), flush=True)
        torch.cuda.empty_cache()
PY
echo "wrn rc=$?"; cut -c1-260 gpurun_out/r3o_wrn.log | tail -6
