set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
python tools/tapgemm_breakdown.py > gpurun_out/breakdown_a.json 2> gpurun_out/breakdown_a.err; tail -3 gpurun_out/breakdown_a.err
