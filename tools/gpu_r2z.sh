#!/bin/bash
# full regression after the tcgen05 attention blocks: GPU suite, smoke(), bench, training step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r2z_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['other_kernels'].items()}, round(d['roofline']['kernel_ms_per_step'],3), d['roofline']['frac'], d['clocks'])
P
timeout 300 python tools/bench_configs.py T3 > gpurun_out/r2z_T3.log 2>&1; tail -1 gpurun_out/r2z_T3.log | cut -c1-400
