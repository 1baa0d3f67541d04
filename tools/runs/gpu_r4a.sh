#!/bin/bash
# round 2, step 4a: channels-last MaxPool2d (csrc/pool_nhwc.cu) -- parity, model tests, benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ibn.py -m gpu -q -k "maxpool" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -q -k "resnet50" 2>&1 | tail -3
timeout 1200 python - > gpurun_out/r4a_bench.log 2>&1 <<'PY'
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200 import train
dev = torch.device("cuda", 0)
keys = ("value", "ms_per_step", "memory_format")
r = train.bench_resnet50(dev, 1, 0, steps=8, warmup=3)
print("r50", json.dumps({k: r[k] for k in keys}), flush=True)
torch.cuda.empty_cache()
r = train.bench_resnet50_jsd(dev, 1, 0, steps=5, warmup=3)
print("jsd", json.dumps({k: r[k] for k in keys}), flush=True)
PY
cat gpurun_out/r4a_bench.log | tail -3
