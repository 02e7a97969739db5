#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c_probe.log; : > $L
for c in 0 1 2 3 4 5; do timeout 90 python -u tools/probe_train.py layers $c >> $L 2>&1; echo "rc=$? case $c" >> $L; done
timeout 90 python -u tools/probe_train.py norms >> $L 2>&1; echo "rc=$? norms" >> $L
timeout 240 python -u tools/probe_train.py golden >> $L 2>&1; echo "rc=$? golden" >> $L
grep -v "^$" $L | grep -v Warning | tail -60
timeout 120 python tools/diag_batch_invariance.py > gpurun_out/r2c_diag.log 2>&1; head -50 gpurun_out/r2c_diag.log
timeout 200 python tools/bench_r1_loop.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"r1-loop\", d[\"value\"], d[\"ms_per_step\"], d[\"e2e\"][\"value\"])"; timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"r2-bench\", d[\"value\"], d[\"ms_per_step\"], d[\"e2e\"][\"value\"])"
