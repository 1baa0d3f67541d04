#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; tail -6 gpurun_out/r2a_pytest.log
