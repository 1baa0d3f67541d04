#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/s18_bench.log 2>&1
tail -5 gpurun_out/s18_bench.log | cut -c1-3000
( time timeout 600 python bench.py --impl reference ) > gpurun_out/s18_ref.log 2>&1
tail -5 gpurun_out/s18_ref.log | cut -c1-1500
