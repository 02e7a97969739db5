python -m pytest tests/test_gpu_wavelets.py tests/test_coef_builders.py tests/test_gpu_pipeline.py tests/test_gpu_burgers.py -q -m gpu --maxfail=6 2>&1 | tail -25
python tools/bench_dwt.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print('   %-62s %7.1f us  %6.0f GB/s' % (d['transform'], d['ms'] * 1e3, d['algorithmic_GBps']))"
