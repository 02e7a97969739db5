"""N>1 host logic on CPU (gloo, world_size 2): batch sharding + the single all-gather of final fields."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wdno_b200 import parallel as P


def test_shard_bounds_cover_batch():
    for batch in (1, 2, 7, 16, 128):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeDiffusion:
    """stands in for GaussianDiffusion.sample(): deterministic function of the conditions"""

    def sample(self, batch_size, init=None, control=None, **kw):
        assert init.shape[0] == batch_size and control.shape[0] == batch_size
        return init[:, None] * 2.0 + control.sum(dim=2)


class _FakeNoisyDiffusion:
    """sample() consumes noise the way the engine classes do: through `_noise_source` when set, else torch.randn"""

    def __init__(self):
        self._noise_source = None

    def _randn(self, shape, device):
        if self._noise_source is not None:
            return self._noise_source(tuple(shape), device)
        return torch.randn(shape, device=device)

    def sample(self, batch_size, init=None, **kw):
        x = self._randn((batch_size, 3, 4), init.device)
        for _ in range(3):
            x = 0.5 * x + init + self._randn((batch_size, 3, 4), init.device)
        return x


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    init = torch.randn(batch, 3, 4, generator=g)
    control = torch.randn(batch, 1, 5, 3, 4, generator=g)
    out = P.sample_sharded(_FakeDiffusion(), batch, post=lambda x: x + 1.0, init=init, control=control)
    full = _FakeDiffusion().sample(batch, init=init, control=control) + 1.0
    noise = P.full_batch_noise((batch, 2), world, rank, generator=torch.Generator().manual_seed(5))
    ref = torch.randn((batch, 2), generator=torch.Generator().manual_seed(5))
    lo, hi = P.shard_bounds(batch, world, rank)
    # RNG rule: same seed on every rank + full-batch draws sliced per rank == the single-process trajectory
    torch.manual_seed(7)
    noisy = P.sample_sharded(_FakeNoisyDiffusion(), batch, init=init)
    torch.manual_seed(7)
    single = _FakeNoisyDiffusion().sample(batch, init=init)
    ok_rng = bool(torch.equal(noisy, single))
    q.put((rank, bool(out.shape == full.shape and torch.allclose(out, full, atol=1e-6)),
           bool(torch.equal(noise, ref[lo:hi])) and ok_rng))
    dist.destroy_process_group()


def test_sample_sharded_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    for batch in (4, 5):  # even and ragged split
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert all(ok1 and ok2 for _, ok1, ok2 in res), res
