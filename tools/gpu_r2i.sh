#!/bin/bash
# knob sweep, parked A/B measurements (closed-form guidance, WDNO_DWT2D_V2), C2/C4/C5 bench lines
mkdir -p gpurun_out
L=gpurun_out/r2i.log; : > $L
timeout 900 python -m pytest tests/test_gpu_knobs.py -q --timeout 300 2>&1 | grep -v Warning | tail -15 >> $L; echo "rc=$? knobs" >> $L
WDNO_TEST_EXPERIMENTAL=1 WDNO_DWT2D_V2=1 timeout 300 python -m pytest tests/test_gpu_wavelets.py -q -k v2 --timeout 200 2>&1 | tail -5 >> $L; echo "rc=$? dwt2d_v2 test" >> $L
timeout 200 python tools/bench_dwt.py >> gpurun_out/r2i_dwt_default.jsonl 2>> $L; WDNO_DWT2D_V2=1 timeout 200 python tools/bench_dwt.py >> gpurun_out/r2i_dwt_v2.jsonl 2>> $L
grep -h "graph" gpurun_out/r2i_dwt_default.jsonl | cut -c1-300 >> $L; echo "--- v2" >> $L; grep -h "graph" gpurun_out/r2i_dwt_v2.jsonl | cut -c1-300 >> $L
timeout 600 python tools/bench_configs.py C5 C2 C4 --steps 20 > gpurun_out/r2i_configs.jsonl 2>> $L; cut -c1-1200 gpurun_out/r2i_configs.jsonl >> $L
grep -v "^$" $L | tail -60
