#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -k "wrn40_2" 2>&1 | tail -6
timeout 900 python - > gpurun_out/r3s_wrn.log 2>&1 <<'PY'
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200 import train
dev = torch.device("cuda", 0)
for cl in (True, False, True):
    r = train.bench_wrn(dev, 1, 0, steps=40, warmup=8, cn_prob=0.25, fuse_post=True, channels_last=cl)
    print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "final_loss", "memory_format", "cnsn_kernel_launches")}), flush=True)
    torch.cuda.empty_cache()
PY
echo "wrn rc=$?"; cut -c1-260 gpurun_out/r3s_wrn.log | tail -6
