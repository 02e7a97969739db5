# end-of-round evidence run (one GPU): tests, ncu captures, per-config profiles and benches -> gpurun_out/ (then tools/make_profiles.py r1d)
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_prof_r1d.sh
python tools/step_profile.py 2>&1 | tail -13 > gpurun_out/step_profile_C3.txt
CONFIG=C2 python tools/step_profile.py 2>&1 | tail -12 > gpurun_out/step_profile_C2.txt
CONFIG=C4 python tools/step_profile.py 2>&1 | tail -12 > gpurun_out/step_profile_C4.txt
python tools/tapgemm_breakdown.py > /dev/null 2>&1
CONFIG=C2 python tools/tapgemm_breakdown.py > /dev/null 2>&1
CONFIG=C4 python tools/tapgemm_breakdown.py > /dev/null 2>&1
python tools/bench_configs.py C2 C4 C5 > gpurun_out/bench_configs.jsonl 2>/dev/null
python tools/bench_dwt.py > gpurun_out/bench_dwt.jsonl 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"dwt|3d_kernel" -c 4 --csv --log-file gpurun_out/dwt_ncu.csv python tools/bench_dwt.py > /dev/null 2>&1
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
cut -c1-330 gpurun_out/bench_r1d.json
cut -c1-200 gpurun_out/bench_configs.jsonl
