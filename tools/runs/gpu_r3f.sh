#!/bin/bash
# 8-GPU check of the graphed ResNet-50 / JSD steps (flat-buffer all-reduce instead of DDP)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-crossnorm --no-sustained > gpurun_out/r3f_bench8.log 2> gpurun_out/r3f_bench8.err
echo "== rc=$?"; head -c 700 gpurun_out/r3f_bench8.log; echo; grep -v "Warning\|run_backward\|^\*\|OMP_NUM\|^$" gpurun_out/r3f_bench8.err | tail -5 | cut -c1-300
