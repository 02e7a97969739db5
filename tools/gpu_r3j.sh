#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 2 -c 1 -f -o gpurun_out/prof_r3j_tapgemm_c64 python tools/tapgemm_once.py 64 64 40 > gpurun_out/r3j_a.log 2>&1; tail -1 gpurun_out/r3j_a.log | cut -c1-80
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 2 -c 1 -f -o gpurun_out/prof_r3j_tapgemm_c256 python tools/tapgemm_once.py 256 256 10 > gpurun_out/r3j_b.log 2>&1; tail -1 gpurun_out/r3j_b.log | cut -c1-80
