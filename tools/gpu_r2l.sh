#!/bin/bash
# regression: full 1-GPU suite + smoke() + bench (C3) after the round-2 kernel changes
mkdir -p gpurun_out
L=gpurun_out/r2l.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v Warning | tail -25 >> $L; echo "rc=$? full gpu suite" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "rc=$? smoke" >> $L
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2>> $L; python - <<'PY' >> $L
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print("bench", d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d.get('cpu_baseline',{}).get('value'))
for k,v in d['roofline']['other_kernels'].items(): print(k, round(v['ms_per_step'],4), round(v.get('frac',0),4))
print('tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])
PY
grep -v "^$" $L | tail -40
