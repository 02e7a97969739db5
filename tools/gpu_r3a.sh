#!/bin/bash
# CTA-pair weight multicast in the tap-GEMM: unit tests under a short timeout first, then the C3 step with and without it
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_tapgemm.py -x -q > gpurun_out/r3a_t1.log 2>&1; echo "tapgemm tests rc=$?"; tail -3 gpurun_out/r3a_t1.log
timeout 300 python -m pytest tests/test_gpu_smoke.py -x -q > gpurun_out/r3a_t2.log 2>&1; echo "smoke tests rc=$?"; tail -3 gpurun_out/r3a_t2.log
for mode in 1 0; do
WDNO_CLUSTER=$mode timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r3a_bench_$mode.json 2> gpurun_out/r3a_bench_$mode.err; echo "bench cluster=$mode rc=$?"
done
python - <<'P'
import json
for m in (1,0):
    for l in open(f'gpurun_out/r3a_bench_{m}.json'):
        if l.startswith('{'):
            d=json.loads(l); print('cluster',m, round(d['value'],2), round(d['ms_per_step'],3), 'tapgemm ms', round(d['roofline']['kernel_ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))
P
