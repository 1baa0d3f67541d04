#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.log 2> gpurun_out/r2k_bench.err ) 2>&1 | tail -3
head -c 1500 gpurun_out/r2k_bench.log; echo; tail -5 gpurun_out/r2k_bench.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_ref.log 2>&1 ) 2>&1 | tail -3
head -c 600 gpurun_out/r2k_ref.log
