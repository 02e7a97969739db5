# streaming 3-D DWT kernels: parity tests, then GPU time per tile configuration (CUDA-graph replays)
python -m pytest tests/test_gpu_wavelets.py tests/test_coef_builders.py tests/test_gpu_pipeline.py tests/test_gpu_smoke.py -x -q -m gpu 2>&1 | tail -5
run() { echo "== $*"; env "$@" python tools/bench_dwt.py --graph3d 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print('   %-58s %7.1f us  %6.0f GB/s' % (d['transform'], d['ms'] * 1e3, d['algorithmic_GBps']))"; }
run WDNO_X=default
run WDNO_DWT3D_ANR=1
run WDNO_DWT3D_ATHREADS=320
run WDNO_DWT3D_ATHREADS=192
run WDNO_DWT3D_ATH=12 WDNO_DWT3D_ATD=18 WDNO_DWT3D_STH=16 WDNO_DWT3D_STD=16
run WDNO_DWT3D_ATH=12 WDNO_DWT3D_ATD=6 WDNO_DWT3D_STH=8 WDNO_DWT3D_STD=4
run WDNO_DWT3D_ATH=18 WDNO_DWT3D_ATD=9 WDNO_DWT3D_STH=16 WDNO_DWT3D_STD=8
run WDNO_DWT3D_ATH=6 WDNO_DWT3D_ATD=18 WDNO_DWT3D_STH=4 WDNO_DWT3D_STD=16
run WDNO_DWT3D_ATH=8 WDNO_DWT3D_ATD=9 WDNO_DWT3D_STH=11 WDNO_DWT3D_STD=8
