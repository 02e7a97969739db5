python tools/bench_dwt.py 2>&1 | tail -7 | cut -c1-175
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"3d_" -c 12 --csv --log-file gpurun_out/dwt_ncu_r1e.csv python tools/bench_dwt.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ana3d_stream" -c 1 -o gpurun_out/prof_ana3d_stream_r1e -f python tools/bench_dwt.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"syn3d_stream" -c 1 -o gpurun_out/prof_syn3d_stream_r1e -f python tools/bench_dwt.py > /dev/null 2>&1
ls -la gpurun_out | tail -5
