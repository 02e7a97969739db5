#!/bin/bash
# round 2, call B (2 GPUs): remaining GPU suite, sharded NCCL test, bench at N=2 and N=1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1500 --deselect tests/test_gpu_parity_configs.py::test_c3_full_ddim250_trajectory_engine_vs_fp32_oracle --deselect tests/test_gpu_parity_configs.py::test_c2_full_ddpm1000_trajectory_engine_vs_fp32_oracle 2>&1 | tail -30 > gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -c 600 gpurun_out/r2b_bench_n2.json; tail -3 gpurun_out/r2b_bench_n2.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.load(open(f'gpurun_out/r2b_bench_n{n}.json'))
        print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    except Exception as e: print(n, 'ERR', e)
PY
