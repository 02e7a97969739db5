#!/bin/bash
# ncu: attention forward kernels in the C3 step (round 2), launch list of one graph-replayed step
mkdir -p gpurun_out
N=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2_step.csv python tools/step_once.py > gpurun_out/r2k_step.log 2>&1; tail -1 gpurun_out/r2k_step.log
for k in tattn_tc_kernel tattn_kernel la1_kernel la2_warp_kernel la_mid_kernel; do
N=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -f -o gpurun_out/prof_r2_$k python tools/step_once.py > gpurun_out/r2k_$k.log 2>&1; tail -1 gpurun_out/r2k_$k.log
done
WDNO_TATTN_TC=0 N=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tattn_warp_kernel -s 0 -c 1 -f -o gpurun_out/prof_r2_tattn_warp_kernel python tools/step_once.py > gpurun_out/r2k_tw.log 2>&1; tail -1 gpurun_out/r2k_tw.log
ls -la gpurun_out/prof_r2_*.ncu-rep
