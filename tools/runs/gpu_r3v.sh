#!/bin/bash
# round 2, step v: full GPU suite, smoke, default bench.py
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r3v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3v_pytest.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r3v_bench.log 2> gpurun_out/r3v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r3v_bench.log") if x.startswith("{")][-1]
d=json.loads(l)
print({k:d[k] for k in ("value","ms_per_step","train_summary","gpu_launches")})
print(d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"])
PY
