#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/s19_bench2.log 2>&1
tail -4 gpurun_out/s19_bench2.log | cut -c1-2500
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/s19_ref2.log 2>&1
tail -2 gpurun_out/s19_ref2.log | cut -c1-600
