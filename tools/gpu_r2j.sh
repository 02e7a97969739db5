#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2j.log; : > $L
timeout 300 python -m pytest tests/test_gpu_attn_fused.py -q -x --timeout 120 2>&1 | grep -v Warning | tail -25 >> $L; echo "rc=$? attn_fused tests" >> $L
timeout 300 python -m pytest tests/test_gpu_smoke.py -q -x --timeout 200 -k "unet3d_forward or ddim_sample" 2>&1 | grep -v Warning | tail -15 >> $L; echo "rc=$? smoke golden" >> $L
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench.json 2>> $L; python - <<'PY' >> $L
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print("bench", d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in d['roofline']['other_kernels'].items(): print(k, round(v['ms_per_step'],4), round(v.get('frac',0),4))
print('tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])
PY
WDNO_TATTN_TC=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('TC=0 bench', d['value'], d['ms_per_step'])" >> $L
grep -v "^$" $L | tail -40
