#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_training.py -q -x --timeout 1500 -s 2>&1 | tail -40 > gpurun_out/r2c_train.log; tail -25 gpurun_out/r2c_train.log
python tools/diag_batch_invariance.py > gpurun_out/r2c_diag.log 2>&1; head -50 gpurun_out/r2c_diag.log
python tools/bench_r1_loop.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"r1-loop\", d[\"value\"], d[\"ms_per_step\"], d[\"e2e\"][\"value\"])"; python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"r2-bench\", d[\"value\"], d[\"ms_per_step\"], d[\"e2e\"][\"value\"])"
