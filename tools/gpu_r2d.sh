#!/bin/bash
# 2 GPUs: training tests, DDP + sharded NCCL tests, training bench
mkdir -p gpurun_out
L=gpurun_out/r2d.log; : > $L
timeout 120 python -u tools/probe_train.py norms >> $L 2>&1; echo "rc=$? norms" >> $L
timeout 600 python -m pytest tests/test_gpu_training.py -q -x --timeout 280 2>&1 | tail -30 >> $L; echo "rc=$? training tests" >> $L
timeout 600 python -m pytest tests/test_gpu_sharded.py -q --timeout 500 2>&1 | tail -30 >> $L; echo "rc=$? sharded+ddp" >> $L
timeout 300 python tools/bench_configs.py T3 --steps 10 >> $L 2>&1; echo "rc=$? T3" >> $L
grep -v "^$" $L | grep -v Warning | tail -70
