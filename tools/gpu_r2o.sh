#!/bin/bash
# bench.py --config C4 / C5 / C2 on one GPU (first run of those code paths)
mkdir -p gpurun_out
for c in C4 C5 C2; do
timeout 400 python bench.py --config $c --steps 10 --warmup 3 --e2e-steps 20 --no-cpu-baseline > gpurun_out/r2o_$c.json 2> gpurun_out/r2o_$c.err; echo "rc=$? $c"; tail -2 gpurun_out/r2o_$c.err | cut -c1-400
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2o_$c.json'))
    print("$c", d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['whole_step_frac_of_sustained_peak'])
except Exception as e: print("$c ERR", e)
PY
done
