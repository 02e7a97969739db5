#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2u.log
for c in tiny one small64 small128 split64 split128 ragged full64 full128; do
  timeout 90 python tools/probe_linattn_tc.py $c >> gpurun_out/r2u.log 2>&1
  echo "rc=$? $c" >> gpurun_out/r2u.log
done
grep -v "^rc=0" gpurun_out/r2u.log | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches.csv python tools/probe_linattn_tc.py full64 full128 > gpurun_out/r2u_probe.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2u_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    k=r[ki][:60]
    if 'wdno' in k: agg.setdefault(k,[]).append(float(r[vi].replace(',','')))
for k,v in agg.items():
    print(f"{k:62s} n={len(v):3d} median={sorted(v)[len(v)//2]/1e3:8.1f} us")
P
