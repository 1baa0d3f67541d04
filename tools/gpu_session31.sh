#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
( timeout 600 python bench.py --no-train ) 2>&1 | tail -1 | cut -c1-300
