#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_graph_gpu.py -x -q -m gpu 2>&1 | grep -v Warning | tail -4
timeout 600 python - <<'EOF' 2>&1 | grep -v "Warning\|run_backward" | tee gpurun_out/r3a_r50.log
import sys, json; sys.path.insert(0, '.')
import torch
from cnsn_b200.train import bench_resnet50, bench_resnet50_jsd
dev = torch.device('cuda', 0)
for graph in (False, True):
    r = bench_resnet50(dev, 1, 0, batch=256, steps=8, warmup=3, fuse_post=True, graph=graph)
    print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'final_loss', 'graph')}))
    torch.cuda.empty_cache()
for graph in (False, True):
    r = bench_resnet50_jsd(dev, 1, 0, batch=256, steps=4, warmup=2, fuse_post=True, graph=graph)
    print(json.dumps({k: r[k] for k in ('value', 'ms_per_step', 'final_loss', 'peak_mem_gb')}))
    torch.cuda.empty_cache()
EOF
