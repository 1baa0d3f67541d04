#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ibn.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/perf_bn.py 2>&1 | tee gpurun_out/r2x_bn.log | tail -4
timeout 600 python - <<'EOF' 2>&1 | grep -v Warning | tee gpurun_out/r2x_r50.log
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200.train import bench_resnet50, bench_resnet50_jsd
dev = torch.device('cuda', 0)
r = bench_resnet50(dev, 1, 0, batch=256, steps=8, warmup=3, fuse_post=True)
print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'final_loss', 'cnsn_kernel_launches')}))
r = bench_resnet50_jsd(dev, 1, 0, batch=256, steps=4, warmup=2, fuse_post=True)
print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'final_loss', 'cnsn_kernel_launches')}))
EOF
