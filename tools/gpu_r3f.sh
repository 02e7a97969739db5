#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_smoke.py -x -q 2>&1 | tail -2
timeout 120 python - <<'P'
import os, subprocess, sys
sys.path.insert(0, 'tools')
src = open('tools/sweep_zstack.py').read()
code = src[src.index("CODE = r'''") + len("CODE = r'''"):src.index("''' % ROOT")] % os.getcwd()
for shape in ("64,64,40", "128,64,40", "256,256,10"):
    for act in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SHAPE=shape, ACT=act), capture_output=True, text=True, timeout=100)
        print([l for l in r.stdout.splitlines() if l.startswith("RES")] or r.stderr[-300:])
P
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r3f_bench.json')); print(d['value'], d['ms_per_step'], 'tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
