#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "edge_shapes" 2>&1 | tail -30) > gpurun_out/s29_pytest.log
tail -30 gpurun_out/s29_pytest.log | cut -c1-200
