#!/bin/bash
# DRAM traffic of the tap-GEMM / conv1x1 launches of one C3 step and the launch list, with the final planner
mkdir -p gpurun_out
N=1 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"tapgemm|conv1x1" -c 44 --csv --log-file gpurun_out/tapgemm_traffic_r2.csv python tools/step_once.py > gpurun_out/r3o_a.log 2>&1; echo "rc=$?"
N=2 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2_step.csv python tools/step_once.py > gpurun_out/r3o_b.log 2>&1; echo "rc=$?"
