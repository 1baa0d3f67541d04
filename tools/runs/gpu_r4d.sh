#!/bin/bash
# round 2, step 4d: fused bottleneck tail (bn3 -> add -> SelfNorm -> ReLU), channels-last
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ibn.py -m gpu -q -k "bottleneck_tail" > gpurun_out/r4d_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r4d_tests.log
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -k "resnet50" 2>&1 | tail -3
timeout 1200 python - > gpurun_out/r4d_bench.log 2>&1 <<'PY'
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200 import train
import cnsn_b200.hosts._norm as HN
dev = torch.device("cuda", 0)
keys = ("value", "ms_per_step", "memory_format")
for fused in (False, True):
    HN.FUSE_TAIL = fused
    r = train.bench_resnet50(dev, 1, 0, steps=8, warmup=3)
    print("r50 fused_tail=%s" % fused, json.dumps({k: r[k] for k in keys}), flush=True)
    torch.cuda.empty_cache()
    r = train.bench_resnet50_jsd(dev, 1, 0, steps=5, warmup=3)
    print("jsd fused_tail=%s" % fused, json.dumps({k: r[k] for k in keys}), flush=True)
    torch.cuda.empty_cache()
PY
cat gpurun_out/r4d_bench.log | tail -4
