# streaming 3-D DWT kernels: parity tests, then throughput per tile configuration
python -m pytest tests/test_gpu_wavelets.py tests/test_coef_builders.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -5
echo "== default"; python tools/bench_dwt.py 2>&1 | head -3 | cut -c1-140
echo "== tile kernels"; WDNO_DWT3D_STREAM=0 python tools/bench_dwt.py 2>&1 | head -3 | cut -c1-140
for cfg in "17 18 16 16" "9 9 8 8" "9 6 8 6" "6 18 4 16" "17 9 16 8" "12 18 11 16"; do
  set -- $cfg
  echo "== ATH=$1 ATD=$2 STH=$3 STD=$4"
  WDNO_DWT3D_ATH=$1 WDNO_DWT3D_ATD=$2 WDNO_DWT3D_STH=$3 WDNO_DWT3D_STD=$4 python tools/bench_dwt.py 2>&1 | head -3 | cut -c1-140
done
