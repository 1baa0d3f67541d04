#!/bin/bash
mkdir -p gpurun_out
F="CNSN_SELFNORM_IMPL=flow"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sn_flow --csv --log-file gpurun_out/s3_ncu.csv \
  python tools/sweep_selfnorm.py 256,256,56,56 f32 1 "$F" "$F CNSN_FLOW_KEEP=1 CNSN_FLOW_D=6" "$F CNSN_FLOW_TPI=32" "$F CNSN_FLOW_TPI=32 CNSN_FLOW_KEEP=1 CNSN_FLOW_D=6" "$F CNSN_FLOW_TPI=32 CNSN_FLOW_D=12 CNSN_FLOW_ORDER=1" > gpurun_out/s3_sweep.log 2>&1
tail -5 gpurun_out/s3_sweep.log
