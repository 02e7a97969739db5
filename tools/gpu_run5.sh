python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_e.err | cut -c1-250
WDNO_PDL=0 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>> gpurun_out/bench_e.err | cut -c1-250
tail -3 gpurun_out/bench_e.err
