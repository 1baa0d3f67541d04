#!/bin/bash
# round 2, step 4g: ncu details of the fused-tail kernels at a ResNet-50 stage-1 tail shape
mkdir -p gpurun_out
cat > /tmp/tail_probe.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
import cnsn_b200.cnsn as M
import cnsn_b200.hosts._norm as HN
from cnsn_b200.ibn import BatchNorm2d
dev = "cuda:0"; cl = torch.channels_last
shape = (64, 256, 56, 56)
bn = BatchNorm2d(256).to(dev).train(); site = M.CNSN(None, M.SelfNorm(256)).to(dev).train()
c = torch.randn(shape, device=dev).contiguous(memory_format=cl).requires_grad_(True)
sk = torch.randn(shape, device=dev).contiguous(memory_format=cl).requires_grad_(True)
dy = torch.randn(shape, device=dev).contiguous(memory_format=cl)
for _ in range(3):
    y = HN.bn_site_relu(bn, site, c, sk)
    torch.autograd.grad(y, (c, sk), dy)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_nhwc_|k_bn_nhwc_|k_tail_" -s 10 -c 10 -o gpurun_out/r4g_tail python /tmp/tail_probe.py > gpurun_out/r4g_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r4g_tail.ncu-rep --page details --csv > gpurun_out/r4g_tail_ncu_details.csv 2>/dev/null
rm -f gpurun_out/r4g_tail.ncu-rep
wc -l gpurun_out/r4g_tail_ncu_details.csv
