# r1e evidence (one GPU): what the driver runs at round end (GPU suite, smoke(), both bench arms) + DWT throughput, the ncu launch
# list of the wavelet kernels, the guided C5 step.  Full ncu captures of the streaming 3-D kernels: tools/gpu_dwt_stream2.sh.
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; cut -c1-330 gpurun_out/bench_r1e.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
python tools/bench_dwt.py > gpurun_out/bench_dwt_r1e.jsonl 2>/dev/null
python -c "
import json
for l in open('gpurun_out/bench_dwt_r1e.jsonl'):
    d = json.loads(l); print('   %-62s %7.1f us  %6.0f GB/s' % (d['transform'], d['ms'] * 1e3, d['algorithmic_GBps']))"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"3d_|dwt2d" -c 40 --csv --log-file gpurun_out/dwt_ncu_r1e.csv python tools/bench_dwt.py > /dev/null 2>&1
python tools/bench_configs.py C5 > gpurun_out/bench_configs_r1e.jsonl 2>/dev/null; cut -c1-700 gpurun_out/bench_configs_r1e.jsonl
