#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/micro/tmem_ld_rate > gpurun_out/r2s_tmem.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_tmem.log
cat gpurun_out/r2s_tmem.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches.csv python tools/probe_linattn_tc.py full64 full128 > gpurun_out/r2s_probe.log 2>&1
echo "rc=$?"
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2s_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    k=r[ki][:60]; agg.setdefault(k,[]).append(float(r[vi].replace(',','')))
for k,v in agg.items():
    print(f"{k:62s} n={len(v):3d} median={sorted(v)[len(v)//2]/1e3:8.1f} us")
P
