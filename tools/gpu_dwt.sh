# DWT / IDWT throughput (fast paths vs the generic kernels) + ncu DRAM metrics of the 3-D transform's passes
python tools/bench_dwt.py 2>&1 | tee gpurun_out/bench_dwt.jsonl | cut -c1-250
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:dwt -c 6 --csv --log-file gpurun_out/dwt_ncu.csv python tools/bench_dwt.py > /dev/null 2>&1
grep -v "^==" gpurun_out/dwt_ncu.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -30
