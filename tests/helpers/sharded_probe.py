"""torchrun --nproc-per-node 2 worker of tests/test_gpu_sharded.py: the multi-GPU PRODUCT path on NCCL --
`parallel.sample_sharded(diffusion, B, post=<state -> fields>)` with the real engine, full-batch noise sliced per rank,
one all-gather of the `[B,32,6,64,64]` fields (reference call shape smoke/inference_2d.py:123-152)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from wdno_b200 import parallel as P  # noqa: E402
from wdno_b200.diffusion_smoke import GaussianDiffusion  # noqa: E402
from wdno_b200.smoke.inference_2d import state_to_fields  # noqa: E402
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).to(dev).eval()
    rescaler = torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1).to(dev)
    shape, ori = [18, 34, 34], [32, 64, 64]
    gd = GaussianDiffusion(m, rescaler, True, True, True, False, "bior1.3", "zero", shape, ori, image_size=40, frames=24,
                           timesteps=1000, sampling_timesteps=3, ddim_sampling_eta=1.0).to(dev)
    post = lambda x: state_to_fields(x, rescaler, shape, ori, "bior1.3", "zero")
    res = {}
    for B in (4, 5):   # even and ragged split
        g = torch.Generator().manual_seed(1234)
        init = torch.randn(B, 24, 40, 40, generator=g).to(dev)
        control = torch.randn(B, 24, 16, 40, 40, generator=g).to(dev)
        torch.manual_seed(11)
        out = P.sample_sharded(gd, B, post=post, init=init, control=control)
        assert tuple(out.shape) == (B, 32, 6, 64, 64), out.shape
        # every rank holds the same gathered tensor
        chk = out.double().sum().reshape(1)
        lst = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        same_everywhere = all(bool(torch.equal(lst[0], t)) for t in lst)
        if rank == 0:
            # (1) each shard == a single-process run of that shard fed with its rows of the full-batch noise (bit-equal)
            shard_equal = []
            for r in range(world):
                lo, hi = P.shard_bounds(B, world, r)
                torch.manual_seed(11)
                gd._noise_source = P._FullBatchNoise(B, lo, hi)
                exp = post(gd.sample(batch_size=hi - lo, init=init[lo:hi], control=control[lo:hi]))
                gd._noise_source = None
                shard_equal.append(bool(torch.equal(out[lo:hi], exp)))
            # (2) the whole gathered batch vs the single-GPU run of the full batch with the same seed (plans may differ
            #     with the batch size -> different fp32 partial-sum grouping of the GroupNorm statistics: round-off only)
            torch.manual_seed(11)
            full = post(gd.sample(batch_size=B, init=init, control=control))
            rel = float((out.double() - full.double()).norm() / full.double().norm())
            res[str(B)] = dict(shard_equal=shard_equal, rel_l2_vs_single_gpu=rel, bit_equal_single=bool(torch.equal(out, full)),
                               same_everywhere=same_everywhere, finite=bool(torch.isfinite(out).all()))
        dist.barrier()
    if rank == 0:
        print("SHARDED_JSON " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
