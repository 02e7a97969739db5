#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2w.log
for c in tiny small64 split64 split128 ragged full64 full128; do
  timeout 90 python tools/probe_linattn_tc.py $c >> gpurun_out/r2w.log 2>&1
  echo "rc=$? $c" >> gpurun_out/r2w.log
done
grep -v "^rc=0" gpurun_out/r2w.log | cut -c1-420
timeout 300 python -m pytest tests/test_gpu_attn_fused.py -x -q 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r2w_bench.json',):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['other_kernels'].items()}, round(d['roofline']['kernel_ms_per_step'],3))
P
