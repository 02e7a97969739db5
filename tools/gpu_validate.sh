# what the driver runs at round end, on one GPU: GPU tests, smoke(), both bench arms
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
( time python bench.py ) > gpurun_out/bench_validate.json 2> gpurun_out/bench_validate.err; tail -4 gpurun_out/bench_validate.err
cut -c1-400 gpurun_out/bench_validate.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) 2>&1 | tail -5 | cut -c1-600
