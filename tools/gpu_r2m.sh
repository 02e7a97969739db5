#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2m.log; : > $L
for sw in 0 1; do for c in 0 1 2 3 4; do SWAP=$sw CASE=$c timeout 60 python -u tools/probe_wgrad_tc.py >> $L 2>&1; echo "rc=$? swap $sw case $c" >> $L; done; done
grep -v "^$" $L | grep -v Warning | tail -60
