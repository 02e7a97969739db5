#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2p.log; : > $L
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_training.py -q -x --timeout 250 2>&1 | grep -v Warning | grep -v "^$" | tail -12 >> $L; echo "rc=$? pipeline+training tests" >> $L
timeout 300 python tools/bench_configs.py T3 --steps 5 2>&1 | cut -c1-900 >> $L; echo "rc=$? T3" >> $L
timeout 400 python bench.py --config C5 --steps 20 --warmup 5 --e2e-steps 50 --no-cpu-baseline > gpurun_out/r2p_C5.json 2> gpurun_out/r2p_C5.err; echo "rc=$? C5" >> $L; grep -c "not capturable" gpurun_out/r2p_C5.err >> $L
python - >> $L <<'PY'
import json
d=json.load(open('gpurun_out/r2p_C5.json')); print("C5", d['value'], d['ms_per_step'], d['e2e']['value'])
PY
grep -v "^$" $L | tail -30
