#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2h.log; : > $L
timeout 500 python -m pytest tests/test_gpu_training.py -q -x --timeout 200 -s -k "burgers or wgrad_and_dgrad" 2>&1 | grep -v Warning | tail -25 >> $L; echo "rc=$? burgers/wgrad tests" >> $L
timeout 300 python tools/bench_configs.py T3 --steps 5 >> $L 2>&1; echo "rc=$? T3" >> $L
grep -v "^$" $L | tail -40 | cut -c1-1500
