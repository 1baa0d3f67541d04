#!/bin/bash
# 8-GPU run: the driver's scaling command at N = 8 (and N = 4 on the same box)
mkdir -p gpurun_out
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2z_bench$n.log 2> gpurun_out/r2z_bench$n.err
  echo "== N=$n rc=$?"; head -c 600 gpurun_out/r2z_bench$n.log; echo; grep -v "Warning\|run_backward\|^$" gpurun_out/r2z_bench$n.err | tail -3 | cut -c1-300
done
python - <<'EOF'
import json
for n in (8, 4):
    try:
        d = json.loads(open("gpurun_out/r2z_bench%d.log" % n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"], 1), "copy_only GB/s/dir/gpu", round(d["e2e"]["copy_only"]["gbs_per_direction_per_gpu"], 1), d["train_summary"])
    except Exception as e:
        print(n, "ERR", e)
EOF
