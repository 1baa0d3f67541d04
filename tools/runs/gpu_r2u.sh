#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest.log 2>&1; tail -5 gpurun_out/r2u_pytest.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench.log 2> gpurun_out/r2u_bench.err; head -c 700 gpurun_out/r2u_bench.log; echo
python - <<'EOF'
import json
d = json.loads(open("gpurun_out/r2u_bench.log").read().strip().splitlines()[-1])
print("e2e", json.dumps(d["e2e"])[:700])
print("crossnorm", json.dumps(d["crossnorm"])[:900])
EOF
