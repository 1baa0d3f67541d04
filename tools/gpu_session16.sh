#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s16_pytest.log
tail -3 gpurun_out/s16_pytest.log
{
for cfg in "128,64,32,32 bf16 both" "512,32,32,32 f32 both" "512,64,16,16 f32 both" "256,64,56,56 f32 both" "256,64,56,56 f32 style" "256,64,56,56 f32 content"; do set -- $cfg
  timeout 120 python tools/perf_cabi.py crossnorm $1 $2 $3 50
done
PERF_EVAL=1 timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 f32 neither 30
PERF_EVAL=1 timeout 120 python tools/perf_cabi.py selfnorm 256,1024,14,14 f32 neither 30
PERF_EVAL=1 CNSN_SELFNORM_IMPL=v1 timeout 120 python tools/perf_cabi.py selfnorm 256,256,56,56 f32 neither 30
} > gpurun_out/s16_perf.log 2>&1
cat gpurun_out/s16_perf.log
timeout 400 python bench.py --no-train > gpurun_out/s16_bench.log 2>&1
tail -1 gpurun_out/s16_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['fwd'], d['roofline']['bwd']); print(json.dumps(d['crossnorm'],indent=1)); print(d['e2e']); print(d['cpu_baseline'])"
