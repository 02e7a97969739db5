#!/bin/bash
# 8 GPUs: the product path (sample_sharded over NCCL) for C3 / C4 (batch 128 over 8) / C5 (batch 64 over 8, guided)
mkdir -p gpurun_out
for c in C3 C4 C5; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config $c --steps 20 --warmup 5 > gpurun_out/r3d_$c.json 2> gpurun_out/r3d_$c.err; echo "rc=$? $c"
python - <<PY
import json
try:
    s=open('gpurun_out/r3d_$c.json').read(); d=json.loads(s[s.index('{"metric'):])
    print("$c N=8", d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_total'], d['e2e']['ms_d2h_of_gathered_fields_rank0'], d['clocks'])
except Exception as e: print("$c ERR", e)
PY
done
timeout 300 python -m pytest tests/test_gpu_sharded.py -q --timeout 280 2>&1 | tail -3
