"""SURVEY.md section 8(e) on hardware: `sample_sharded` with the real engine over NCCL (2 ranks), full-batch noise sliced per
rank, `post` = coefficients -> fields, one all-gather.  Needs 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sample_sharded_nccl_world2_matches_single_gpu():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(HERE, "helpers", "sharded_probe.py")], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("SHARDED_JSON ")][-1]
    res = json.loads(line[len("SHARDED_JSON "):])
    for B, d in res.items():
        assert d["finite"] and d["same_everywhere"], (B, d)
        assert all(d["shard_equal"]), f"batch {B}: a shard differs from the single-process run of that shard: {d}"
        # 3 DDIM steps from t = 999: identical noise, identical kernels; only the batch-dependent plan may regroup fp32 sums
        assert d["rel_l2_vs_single_gpu"] < 1e-3, (B, d)
