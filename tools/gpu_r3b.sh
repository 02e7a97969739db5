#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r3b.log
for c in c2 c3 c256; do
  echo "== $c" >> gpurun_out/r3b.log
  timeout 60 python tools/diag_cluster.py $c >> gpurun_out/r3b.log 2>&1
  echo "rc=$?" >> gpurun_out/r3b.log
done
grep -v "^$" gpurun_out/r3b.log | grep -v "Warning\|Search for\|CUDA kernel errors\|For debugging\|Compile with\|max abs" | cut -c1-200 | tail -40
