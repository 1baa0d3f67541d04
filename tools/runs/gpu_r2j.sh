#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ibn.py tests/test_graph_gpu.py tests/test_models_gpu.py tests/test_binding_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python - <<'EOF' 2>&1 | grep -v Warning | tee gpurun_out/r2j_wrn.log
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200.train import bench_wrn, bench_resnet50
dev = torch.device('cuda', 0)
for graph in (False, True):
    r = bench_wrn(dev, 1, 0, batch=512, steps=40, warmup=8, fuse_post=True, graph=graph)
    print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'graph', 'final_loss', 'cnsn_kernel_launches')}))
r = bench_resnet50(dev, 1, 0, batch=256, steps=8, warmup=3, fuse_post=True)
print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'final_loss', 'cnsn_kernel_launches')}))
EOF
timeout 300 python tools/debug/wrn_profile.py benchmark 2>&1 | cut -c1-200 | tee gpurun_out/r2j_prof.log | head -34
