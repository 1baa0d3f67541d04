#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "channels_last" 2>&1 | tail -4
timeout 600 python tools/debug/wrn_graph_probe.py > gpurun_out/r3r_graph.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r3r_graph.log | tail -14
