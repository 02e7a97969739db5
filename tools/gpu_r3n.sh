#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_smoke.py -x -q 2>&1 | tail -1
timeout 100 python tools/sweep_generic.py 128,128,20 2>&1 | head -1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r3n_bench.json')); print('C3', d['value'], d['ms_per_step'], 'tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
timeout 400 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 2> /dev/null | grep '^{' > gpurun_out/r3n_C4.json; python -c "
import json
d=json.load(open('gpurun_out/r3n_C4.json')); print('C4', d['value'], d['ms_per_step'], 'tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
