#!/bin/bash
# producer-side proxy fence of the fused-activation slabs: on (default) vs off -- parity tests and timing
mkdir -p gpurun_out
for pf in 1 0; do
echo "== WDNO_PRODFENCE=$pf"
WDNO_PRODFENCE=$pf timeout 300 python -m pytest tests/test_gpu_smoke.py -x -q 2>&1 | tail -1
WDNO_PRODFENCE=$pf timeout 120 python - <<'P'
import os, subprocess, sys
src = open('tools/sweep_zstack.py').read()
code = src[src.index("CODE = r'''") + len("CODE = r'''"):src.index("''' % ROOT")] % os.getcwd()
for shape in ("64,64,40", "128,64,40", "256,256,10"):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SHAPE=shape, ACT="1"), capture_output=True, text=True, timeout=100)
    print([l for l in r.stdout.splitlines() if l.startswith("RES")] or r.stderr[-300:])
P
WDNO_PRODFENCE=$pf timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r3g_bench_$pf.json 2> gpurun_out/r3g_bench_$pf.err; echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r3g_bench_$pf.json')); print(d['value'], d['ms_per_step'], 'tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
done
