#!/bin/bash
# tcgen05 linear attention: parity/timing of every probe case, then ncu --set full of the two kernels (full-resolution C = 64)
mkdir -p gpurun_out
: > gpurun_out/r2t.log
for c in tiny one small64 small128 split64 split128 ragged full64 full128; do
  timeout 90 python tools/probe_linattn_tc.py $c >> gpurun_out/r2t.log 2>&1
  echo "rc=$? $c" >> gpurun_out/r2t.log
done
grep -v "^rc=0" gpurun_out/r2t.log | cut -c1-400
for k in la1_tc_kernel la2_tc_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r2_$k python tools/probe_linattn_tc.py full64 > gpurun_out/r2t_$k.log 2>&1; tail -1 gpurun_out/r2t_$k.log | cut -c1-200
done
ls -la gpurun_out/prof_r2_la*_tc_kernel.ncu-rep
