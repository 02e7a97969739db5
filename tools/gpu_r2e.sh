#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2e.log; : > $L
timeout 400 python -m pytest tests/test_gpu_training.py -q -x --timeout 280 -k fused 2>&1 | tail -15 >> $L; echo "rc=$? fused test" >> $L
timeout 300 python tools/bench_configs.py T3 --steps 5 >> $L 2>&1; echo "rc=$? T3" >> $L
grep -v "^$" $L | grep -v Warning | tail -40
