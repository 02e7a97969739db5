#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2x.log
for c in tiny small odd f32 mid full; do
  timeout 90 python tools/probe_tattn_row.py $c >> gpurun_out/r2x.log 2>&1
  echo "rc=$? $c" >> gpurun_out/r2x.log
done
grep -v "^rc=0" gpurun_out/r2x.log | cut -c1-420
timeout 300 python -m pytest tests/test_gpu_attn_fused.py -x -q 2>&1 | tail -3
