#!/bin/bash
# ncu evidence, round 2: attention forward kernels (C3 step), wgrad / attention backward (training step), launch lists
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py T3 --steps 5 > gpurun_out/r2g_T3.log 2>&1; tail -2 gpurun_out/r2g_T3.log | cut -c1-900
N=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2_step.csv python tools/step_once.py > gpurun_out/r2g_step.log 2>&1
B=6 N=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2_train.csv python tools/train_once.py > gpurun_out/r2g_train.log 2>&1
for k in tattn_warp_kernel tattn_kernel la1_kernel la2_warp_kernel la_mid_kernel; do
N=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -f -o gpurun_out/prof_r2_$k python tools/step_once.py > gpurun_out/r2g_$k.log 2>&1
done
B=6 N=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_wgrad python tools/train_once.py > gpurun_out/r2g_wgrad.log 2>&1
B=6 N=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:short_attn_bwd -s 0 -c 1 -f -o gpurun_out/prof_r2_short_attn_bwd python tools/train_once.py > gpurun_out/r2g_sab.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -12
