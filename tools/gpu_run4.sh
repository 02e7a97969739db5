python -m pytest tests/test_gpu_attn_fused.py -x -q 2>&1 | tail -5
python tools/step_profile.py 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -3 gpurun_out/bench_d.err; cut -c1-330 gpurun_out/bench_d.json
