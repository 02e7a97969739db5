#!/bin/bash
# tcgen05 linear attention v2: tests, C3 bench (default and WDNO_LINATTN_TC=0), ncu --set full of la1_tc / la2_tc / la_mid
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attn_fused.py -x -q > gpurun_out/r2v_tests.log 2>&1; tail -2 gpurun_out/r2v_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "rc=$?"
WDNO_LINATTN_TC=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 20 > gpurun_out/r2v_bench_old.json 2> gpurun_out/r2v_bench_old.err; echo "rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r2v_bench.json','gpurun_out/r2v_bench_old.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['other_kernels'].items()}, round(d['roofline']['kernel_ms_per_step'],3))
P
for k in la1_tc_kernel la2_tc_kernel la_mid_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r2_$k python tools/probe_linattn_tc.py full64 > gpurun_out/r2v_$k.log 2>&1; tail -1 gpurun_out/r2v_$k.log | cut -c1-100
done
