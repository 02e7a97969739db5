python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -8
python tools/bench_configs.py C5 2>&1 | tail -3 | cut -c1-900
