#!/bin/bash
# compute-sanitizer memcheck over the new tcgen05 attention kernels (small cases)
mkdir -p gpurun_out
for c in tiny split64; do
  timeout 280 compute-sanitizer --tool memcheck --print-limit 5 python tools/probe_linattn_tc.py $c > gpurun_out/r3h_la_$c.log 2>&1; echo "rc=$? linattn $c"; grep -E "ERROR SUMMARY|Invalid|misaligned|case" gpurun_out/r3h_la_$c.log | cut -c1-220 | head -6
done
for c in tiny small; do
  timeout 280 compute-sanitizer --tool memcheck --print-limit 5 python tools/probe_tattn_row.py $c > gpurun_out/r3h_ta_$c.log 2>&1; echo "rc=$? tattn $c"; grep -E "ERROR SUMMARY|Invalid|misaligned|case" gpurun_out/r3h_ta_$c.log | cut -c1-220 | head -6
done
