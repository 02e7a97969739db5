#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tattn_row_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_tattn_row_kernel python tools/probe_tattn_row.py full > gpurun_out/r2y.log 2>&1; tail -1 gpurun_out/r2y.log | cut -c1-100
