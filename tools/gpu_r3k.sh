#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_smoke.py tests/test_gpu_burgers.py -x -q 2>&1 | tail -2
for o in taps slots; do
WDNO_FIT_ORDER=$o timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r3k_bench_$o.json 2> gpurun_out/r3k_bench_$o.err; echo "rc=$? $o"
python -c "
import json
d=json.load(open('gpurun_out/r3k_bench_$o.json')); print('$o', d['value'], d['ms_per_step'], 'tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
done
for c in C4 C2; do
for o in taps slots; do
WDNO_FIT_ORDER=$o timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 2> /dev/null | grep '^{' > gpurun_out/r3k_${c}_$o.json; python -c "
import json
d=json.load(open('gpurun_out/r3k_${c}_$o.json')); print('$c $o', d['value'], d['ms_per_step'], 'tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
done
done
