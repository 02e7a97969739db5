#!/bin/bash
# round 2, final evidence on one B200: full GPU suite, smoke(), bench (all configurations + reference arm), step profile,
# ncu launch list of a C3 step and full captures of the new attention kernels, training step
mkdir -p gpurun_out
L=gpurun_out/r3c.log; : > $L
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v Warning | tail -8 >> $L; echo "rc=$? full gpu suite" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1; echo "rc=$? smoke" >> $L
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r3c_bench.json 2>> $L; echo "rc=$? bench C3" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3c_bench_reference.json 2>> $L; echo "rc=$? bench reference" >> $L
: > gpurun_out/r3c_bench_configs.jsonl
for c in C4 C5 C2; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>> $L | grep '^{' >> gpurun_out/r3c_bench_configs.jsonl; echo "rc=$? bench $c" >> $L
done
timeout 300 python tools/bench_configs.py T3 2>> $L | grep '^{' >> gpurun_out/r3c_bench_configs.jsonl; echo "rc=$? T3" >> $L
timeout 300 python tools/step_profile.py > gpurun_out/r3c_step_profile_C3.txt 2>> $L; echo "rc=$? step profile" >> $L
N=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r3c_launches_step.csv python tools/step_once.py >> $L 2>&1; echo "rc=$? launch list" >> $L
for k in la1_tc_kernel la2_tc_kernel la_mid_kernel tattn_row_kernel; do
  N=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -f -o gpurun_out/prof_r3c_$k python tools/step_once.py >> $L 2>&1; echo "rc=$? ncu $k" >> $L
done
python - <<'PY' >> $L
import json
for f in ('gpurun_out/r3c_bench.json', 'gpurun_out/r3c_bench_reference.json'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, d.get('impl'), d['value'], d['ms_per_step'], d.get('e2e', {}).get('value'), d.get('clocks'), d.get('cpu_baseline', {}).get('value'))
            if 'roofline' in d:
                for k, v in d['roofline'].get('other_kernels', {}).items():
                    print('  ', k, round(v['ms_per_step'], 4), round(v.get('frac', 0), 4))
                print('   tapgemm', d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])
for l in open('gpurun_out/r3c_bench_configs.jsonl'):
    d = json.loads(l)
    print(d.get('config', {}).get('name') if isinstance(d.get('config'), dict) else d.get('config'), d.get('value', d.get('steps_per_s')), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
PY
grep -v "^$" $L | grep -v "==PROF==\|^==WARNING" | tail -45
head -14 gpurun_out/r3c_step_profile_C3.txt
