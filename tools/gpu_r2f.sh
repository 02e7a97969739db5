#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2f.log; : > $L
timeout 300 python -m pytest tests/test_gpu_training.py -q -x --timeout 120 -k "attention_block" 2>&1 | tail -25 >> $L; echo "rc=$? attn block test" >> $L
timeout 500 python -m pytest tests/test_gpu_training.py -q -x --timeout 200 -k "not attention_block" 2>&1 | tail -25 >> $L; echo "rc=$? training tests" >> $L
timeout 300 python tools/bench_configs.py T3 --steps 5 >> $L 2>&1; echo "rc=$? T3" >> $L
grep -v "^$" $L | grep -v Warning | tail -60
