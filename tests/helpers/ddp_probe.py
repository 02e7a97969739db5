"""torchrun --nproc-per-node 2 worker of tests/test_gpu_sharded.py::test_ddp_gradients...: data-parallel training step on NCCL.
Each rank runs p_losses on its sample; FusedTrainer.backward all-reduces the flat gradient in buckets launched from inside the
backward pass.  Check: averaged gradient == mean of the two single-rank gradients (computed locally on rank 0), and both ranks
hold identical parameters after the step."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from wdno_b200.diffusion_smoke import GaussianDiffusion  # noqa: E402
from wdno_b200.trainer import FusedTrainer  # noqa: E402
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: E402


def build(dev):
    torch.manual_seed(0)
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).to(dev).train()
    w = torch.linspace(0.5, 2.0, 42).reshape(1, 1, 42, 1, 1)
    gd = GaussianDiffusion(m, w, True, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64], image_size=40,
                           frames=24, timesteps=1000, sampling_timesteps=250, ddim_sampling_eta=1.0).to(dev)
    return m, gd


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    g = torch.Generator().manual_seed(21)
    x0 = torch.randn(world, 24, 42, 40, 40, generator=g).clamp(-1, 1).to(dev)
    noise = torch.randn(world, 24, 42, 40, 40, generator=g).to(dev)
    t = torch.tensor([100, 800][:world]).to(dev)
    m, gd = build(dev)
    tr = FusedTrainer(gd, lr=1e-4, bucket_mb=8)
    tr.g.zero_()
    loss = gd.p_losses(x0[rank:rank + 1], t[rank:rank + 1], noise[rank:rank + 1].clone())
    tr.backward(loss)
    launched_inside = len(tr._launched)
    g_avg = tr.g.clone()
    tr.optimizer_step()
    # identical parameters on every rank after the step
    chk = tr.p.double().sum().reshape(1)
    lst = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    same = all(bool(torch.equal(lst[0], v)) for v in lst)
    res = None
    if rank == 0:
        m2, gd2 = build(dev)
        from wdno_b200.train3d import flat_grads
        gs = []
        for r in range(world):
            buf = flat_grads(m2)
            buf.zero_()
            gd2.p_losses(x0[r:r + 1], t[r:r + 1], noise[r:r + 1].clone()).backward()
            gs.append(buf.clone())
        want = sum(gs) / world
        rel = float((g_avg.double() - want.double()).norm() / want.double().norm())
        res = dict(rel_grad=rel, bit_equal=bool(torch.equal(g_avg, want)), same_params=same, buckets=len(tr.buckets),
                   finite=bool(torch.isfinite(g_avg).all()))
    dist.barrier()
    if rank == 0:
        print("DDP_JSON " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
