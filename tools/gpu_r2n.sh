#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2n.log; : > $L
timeout 600 python -m pytest tests/test_gpu_training.py -q -x --timeout 250 -s 2>&1 | grep -v Warning | grep -v "^$" | tail -25 >> $L; echo "rc=$? training tests" >> $L
timeout 300 python tools/bench_configs.py T3 --steps 5 >> $L 2>&1; echo "rc=$? T3" >> $L
WDNO_WGRAD_TC=0 timeout 300 python tools/bench_configs.py T3 --steps 5 2>&1 | tail -1 | cut -c1-400 >> $L
grep -v "^$" $L | tail -40 | cut -c1-1600
