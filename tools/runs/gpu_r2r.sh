#!/bin/bash
# 2-GPU run: bench under torch.distributed.run (default and with NUMA binding of the pinned e2e buffers)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2r_bench2.log 2> gpurun_out/r2r_bench2.err; head -c 900 gpurun_out/r2r_bench2.log; echo; tail -3 gpurun_out/r2r_bench2.err | cut -c1-300
CNSN_BENCH_NUMA=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-train --no-crossnorm > gpurun_out/r2r_bench2_numa.log 2> gpurun_out/r2r_bench2_numa.err; python - <<'EOF'
import json
for f in ("gpurun_out/r2r_bench2.log", "gpurun_out/r2r_bench2_numa.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"]["value"], d["e2e"].get("numa_cpus"), "train", d.get("train_summary"))
    except Exception as e:
        print(f, "ERR", e)
EOF
nvidia-smi topo -m | head -12
