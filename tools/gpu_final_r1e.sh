# r1e evidence: full GPU suite, DWT throughput (API-timed and graph-timed), ncu launch list + full captures of the streaming
# 3-D kernels, C5 (guided) step, default bench line
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/bench_dwt.py > gpurun_out/bench_dwt_r1e.jsonl 2>/dev/null; cut -c1-150 gpurun_out/bench_dwt_r1e.jsonl
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"3d_" -c 16 --csv --log-file gpurun_out/dwt_ncu_r1e.csv python tools/bench_dwt.py --graph3d > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ana3d_stream" -c 1 -o gpurun_out/prof_ana3d_stream_r1e -f python tools/bench_dwt.py --graph3d > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"syn3d_stream" -c 1 -o gpurun_out/prof_syn3d_stream_r1e -f python tools/bench_dwt.py --graph3d > /dev/null 2>&1
python tools/bench_configs.py C5 > gpurun_out/bench_configs_r1e.jsonl 2>/dev/null; cut -c1-400 gpurun_out/bench_configs_r1e.jsonl
python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; cut -c1-300 gpurun_out/bench_r1e.json
