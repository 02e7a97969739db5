#!/bin/bash
# round 2: first run of the tcgen05 linear attention.  Every case in its own process under a short timeout.
mkdir -p gpurun_out
: > gpurun_out/r2r.log
for c in tiny one small64 small128 split64 split128 ragged full64 full128; do
  timeout 90 python tools/probe_linattn_tc.py $c >> gpurun_out/r2r.log 2>&1
  echo "rc=$? $c" >> gpurun_out/r2r.log
done
tail -40 gpurun_out/r2r.log
