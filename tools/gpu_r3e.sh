#!/bin/bash
# final regression on one GPU: full suite + smoke() + default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v Warning | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r3e_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches'])"
