python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/tapgemm_breakdown.py > gpurun_out/breakdown_b.txt 2> gpurun_out/breakdown_b.err; tail -3 gpurun_out/breakdown_b.err; head -40 gpurun_out/breakdown_b.txt
