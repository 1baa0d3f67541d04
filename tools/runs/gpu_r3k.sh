#!/bin/bash
# round 2, step k: channel-group BatchNorm kernel (csrc/bn_grp.cu) -- parity, probe against torch, config-5 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ibn.py -m gpu -x -q > gpurun_out/r3k_tests.log 2>&1; echo "tests rc=$?" 
tail -5 gpurun_out/r3k_tests.log
timeout 300 python tools/debug/bn_general_probe.py > gpurun_out/r3k_bnprobe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r3k_bnprobe.log | tail -8
timeout 600 python - > gpurun_out/r3k_jsd.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from cnsn_b200 import train
import json
import torch
dev = torch.device("cuda", 0)
print(json.dumps(train.bench_resnet50_jsd(dev, 1, 0, steps=6, warmup=3)), flush=True)
torch.cuda.empty_cache()
print(json.dumps(train.bench_resnet50(dev, 1, 0, steps=8, warmup=3)), flush=True)
PY
echo "jsd rc=$?"; tail -3 gpurun_out/r3k_jsd.log
