#!/bin/bash
# round 2, call A: full GPU suite (new parity tests included), smoke(), both bench arms
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 1500 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; cat gpurun_out/r2a_bench_ref.json | cut -c1-600
