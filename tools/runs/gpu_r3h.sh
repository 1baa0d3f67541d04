#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r3h_pytest.log 2>&1; tail -4 gpurun_out/r3h_pytest.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 120 compute-sanitizer --tool racecheck --print-limit 5 python - <<'EOF' 2>&1 | tail -4
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import cnsn_b200.cnsn as M, cnsn_b200._lib as L
L.tune(tm_items=0, grid_cap=2)
x = torch.randn(10, 3, 40, 40, device='cuda', requires_grad=True); dy = torch.randn_like(x)
blk = M.CNSN(M.CrossNorm(crop='neither', beta=1), M.SelfNorm(3)).cuda().train()
blk.crossnorm.active = True
y = blk(x); (dx,) = torch.autograd.grad(y, x, dy); torch.cuda.synchronize(); L.async_error(); print('site tmem under racecheck ok', float(dx.abs().sum()))
EOF
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3h_bench.log 2> gpurun_out/r3h_bench.err; head -c 500 gpurun_out/r3h_bench.log; echo
python - <<'EOF'
import json
d = json.loads(open("gpurun_out/r3h_bench.log").read().strip().splitlines()[-1])
print("site", json.dumps(d["site"]))
print("e2e", d["e2e"]["value"], d["e2e"]["frac_of_copy_only"])
EOF
