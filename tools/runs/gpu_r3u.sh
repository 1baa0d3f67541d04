#!/bin/bash
# round 2, step u: channel blocks in the NHWC SelfNorm kernels, CTA-per-channel batch-norm folds; ResNet-50 benches in channels_last
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ibn.py tests/test_gpu_parity.py -m gpu -q -k "channels_last" 2>&1 | tail -4
timeout 1200 python - > gpurun_out/r3u_bench.log 2>&1 <<'PY'
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200 import train
dev = torch.device("cuda", 0)
keys = ("value", "ms_per_step", "final_loss", "memory_format", "cnsn_kernel_launches")
r = train.bench_wrn(dev, 1, 0, steps=40, warmup=8, cn_prob=0.25, fuse_post=True, channels_last=True)
print("wrn", json.dumps({k: r[k] for k in keys}), flush=True)
torch.cuda.empty_cache()
for cl in (False, True):
    r = train.bench_resnet50(dev, 1, 0, steps=8, warmup=3, channels_last=cl)
    print("r50", json.dumps({k: r[k] for k in keys}), flush=True)
    torch.cuda.empty_cache()
for cl in (False, True):
    r = train.bench_resnet50_jsd(dev, 1, 0, steps=5, warmup=3, channels_last=cl)
    print("jsd", json.dumps({k: r[k] for k in keys}), flush=True)
    torch.cuda.empty_cache()
PY
echo "bench rc=$?"; cut -c1-230 gpurun_out/r3u_bench.log | tail -7
