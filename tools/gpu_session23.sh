#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_jsd.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/s23_pytest.log
tail -3 gpurun_out/s23_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/s23_bench.log 2>&1
tail -6 gpurun_out/s23_bench.log | cut -c1-200
python - <<'PY'
import json
for l in open('gpurun_out/s23_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']); print(json.dumps(d['train'], indent=1)[:3000])
PY
