#!/bin/bash
# refresh of the round-end evidence after the planner change: full suite, bench (all configurations), step profile
mkdir -p gpurun_out
L=gpurun_out/r3l.log; : > $L
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v Warning | tail -4 >> $L; echo "rc=$? full gpu suite" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3l_bench.json 2>> $L; echo "rc=$? bench C3" >> $L
: > gpurun_out/r3l_bench_configs.jsonl
for c in C4 C5 C2; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>> $L | grep '^{' >> gpurun_out/r3l_bench_configs.jsonl; echo "rc=$? bench $c" >> $L
done
timeout 300 python tools/bench_configs.py T3 2>> $L | grep '^{' >> gpurun_out/r3l_bench_configs.jsonl; echo "rc=$? T3" >> $L
timeout 300 python tools/step_profile.py > gpurun_out/r3l_step_profile_C3.txt 2>> $L; echo "rc=$? step profile" >> $L
python - <<'PY' >> $L
import json
d = json.load(open('gpurun_out/r3l_bench.json'))
print('C3', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['cpu_baseline']['value'], 'tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['frac_of_burst_peak'], d.get('algorithmic_tflops_whole_step'), d.get('whole_step_frac_of_sustained_peak'))
for k, v in d['roofline']['other_kernels'].items(): print('  ', k, round(v['ms_per_step'], 4), round(v.get('frac', 0), 4))
for l in open('gpurun_out/r3l_bench_configs.jsonl'):
    x = json.loads(l)
    if isinstance(x.get('config'), dict):
        print(x['config']['name'], x['value'], x['ms_per_step'], x['e2e']['value'], x['roofline']['achieved'], x['roofline']['frac'])
    elif 'ms_per_step' in x:
        print(x['config'], x['steps_per_s'], x['ms_per_step'], x.get('samples_per_s'))
PY
grep -v "^$" $L | tail -22; head -8 gpurun_out/r3l_step_profile_C3.txt
