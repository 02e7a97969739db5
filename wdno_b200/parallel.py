"""Batch sharding of independent trajectories over the GPUs of one box.

The path has no data-path collective: every rank denoises its own contiguous slice of the batch with replicated
weights; ONE all-gather of the final (inverse-transformed) fields ends `sample()`.  The reference itself has no
multi-GPU inference (inference_2d.py:550 is single-process); this is the natural extension the north star names.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world_size, rank):
    """contiguous split, first (batch % world) ranks get one extra trajectory"""
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, world_size, rank):
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[0], world_size, rank)
    return t[lo:hi]


def full_batch_noise(shape, world_size, rank, generator=None, device=None):
    """Draw the FULL-batch noise on every rank and slice: a sharded run then consumes exactly the samples the
    single-process run would (Philox is counter based; SURVEY.md section 7 'RNG parity')."""
    full = torch.randn(shape, generator=generator, device=device)
    return shard(full, world_size, rank)


def all_gather_fields(local, batch, group=None):
    """local [b_local, ...] on each rank -> [batch, ...] on every rank (ragged shards padded to the largest)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == mx for lo, hi in sizes):
        # equal shards: one collective straight into the [batch, ...] result, no staging copies
        out = local.new_empty((batch,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat((local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))), 0)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


class _FullBatchNoise:
    """`_noise_source` of a sharded run: every draw is the rank's ROWS of the full-batch tensor its default generator
    would produce (all ranks seeded alike hold the same Philox state) -- the values, and the generator state afterwards,
    are exactly those of the single-process run (SURVEY.md section 7 'RNG parity', section 8e).  On CUDA only the rows
    are evaluated (counter-based Philox: `ops.randn_rows`, csrc/rng.cu); elsewhere the full batch is drawn and sliced."""

    def __init__(self, batch, lo, hi):
        self.batch, self.lo, self.hi = batch, lo, hi

    def __call__(self, shape, device):
        assert shape[0] == self.hi - self.lo, (shape, self.lo, self.hi)
        from . import ops
        return ops.randn_rows(tuple(shape), self.batch, self.lo, device)


def sample_sharded(diffusion, batch_size, post=None, group=None, rng_parity=True, gather=True, **conds):
    """Run diffusion.sample() on this rank's contiguous slice of every batched condition, apply `post` (the inverse
    transform to fields) locally, then ONE all-gather of the result (the only collective of the path).

    conds: the keyword arguments of `sample()`; tensors whose leading dim is the batch are sliced.
    rng_parity=True: noise is drawn at the full-batch shape on every rank and sliced, so that with the same
    `torch.manual_seed` on every rank the sharded run reproduces the single-GPU trajectory sample by sample (on CUDA
    only the rank's rows of that draw are evaluated: no extra traffic).  False: each rank draws only its own shape from its own generator state.
    gather=False returns the local shard (for callers that reduce on their own)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(batch_size, world, rank)
    local = {k: (shard(v, world, rank) if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == batch_size else v)
             for k, v in conds.items()}
    prev = getattr(diffusion, "_noise_source", None)
    swap = rng_parity and world > 1 and prev is None and hasattr(diffusion, "_noise_source")
    if swap:
        diffusion._noise_source = _FullBatchNoise(batch_size, lo, hi)
    try:
        x = diffusion.sample(batch_size=hi - lo, **local)
    finally:
        if swap:
            diffusion._noise_source = prev
    if post is not None:
        x = post(x)
    if not gather:
        return x
    return all_gather_fields(x, batch_size, group)
