#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python - <<'EOF' 2>&1 | tee gpurun_out/r2h_wrn.log
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200.train import bench_wrn
dev = torch.device('cuda', 0)
for graph in (False, True):
    r = bench_wrn(dev, 1, 0, batch=512, steps=40, warmup=8, fuse_post=True, graph=graph)
    print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'graph', 'final_loss', 'cnsn_kernel_launches')}))
EOF
