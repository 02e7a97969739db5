#!/bin/bash
# compute-sanitizer racecheck / synccheck over the new tcgen05 attention kernels (tiny cases)
mkdir -p gpurun_out
for tool in synccheck racecheck; do
  timeout 280 compute-sanitizer --tool $tool --print-limit 5 python tools/probe_linattn_tc.py tiny > gpurun_out/r3i_la_$tool.log 2>&1; echo "rc=$? linattn $tool"; grep -E "SUMMARY|hazard|Barrier error|case" gpurun_out/r3i_la_$tool.log | cut -c1-200 | head -8
  timeout 280 compute-sanitizer --tool $tool --print-limit 5 python tools/probe_tattn_row.py tiny > gpurun_out/r3i_ta_$tool.log 2>&1; echo "rc=$? tattn $tool"; grep -E "SUMMARY|hazard|Barrier error|case" gpurun_out/r3i_ta_$tool.log | cut -c1-200 | head -8
done
