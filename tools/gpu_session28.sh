#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s28_pytest.log
tail -3 gpurun_out/s28_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/s28_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/s28_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value', d['value'], 'frac', d['roofline']['step']['frac'], 'fwd', d['roofline']['fwd']['ms'], 'bwd', d['roofline']['bwd']['ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
        t=d['train']; print('wrn', t['value'], t['ms_per_step']); print('r50', t['resnet50']['value'], t['resnet50']['ms_per_step']); print('jsd', t['resnet50_jsd']['value'], t['resnet50_jsd']['ms_per_step'], t['resnet50_jsd']['peak_mem_gb'])
        print(d['crossnorm'])
PY
tail -4 gpurun_out/s28_bench.log | grep real
