"""One optimiser step on flat memory: backward on the engine -> (bucketed all-reduce over NCCL, overlapped with the rest of
the backward) -> ONE fused clip + Adam + EMA launch.  Same arithmetic as the reference loop
(smoke/ddpm/diffusion_2d.py:1277-1297: `loss.backward()`, `clip_grad_norm_(1.0)`, `Adam.step()`, `MultiStepLR.step()`,
`ema.update()`; DDP through accelerate averages gradients over ranks), without its ~700 per-parameter launches.

The reference's own `Trainer` also works unchanged over the engine classes (their `loss.backward()` fills `p.grad`); this
class is the B200-native loop for the same step."""
import torch
import torch.distributed as dist

from . import _lib
from .train3d import flat_grads


def _p(t):
    import ctypes as C
    return None if t is None else C.c_void_p(t.data_ptr())


def flat_params(model):
    """re-point every trainable parameter's storage into one flat fp32 buffer (values preserved) -> the buffer"""
    params = [p for p in model.parameters() if p.requires_grad]
    buf = getattr(model, "_flat_param", None)
    n = sum(p.numel() for p in params)
    if buf is not None and buf.numel() == n and all(p.data_ptr() == buf.data_ptr() + 4 * o for p, o in zip(params, _offsets(params))):
        return buf
    buf = torch.empty(n, dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        v = buf[off:off + p.numel()].view_as(p)
        v.copy_(p.data)
        p.data = v
        off += p.numel()
    model._flat_param = buf
    return buf


def _offsets(params):
    out, off = [], 0
    for p in params:
        out.append(off)
        off += p.numel()
    return out


def ema_decay(step, beta=0.995, update_after_step=100, inv_gamma=1.0, power=2.0 / 3.0, min_value=0.0):
    """ema_pytorch.EMA.get_current_decay (the package the reference uses, requirements.txt; published warm-up schedule)"""
    epoch = max(step - update_after_step - 1, 0)
    if epoch <= 0:
        return 0.0
    return min(max(1 - (1 + epoch / inv_gamma) ** -power, min_value), beta)


class FusedTrainer:
    def __init__(self, diffusion, lr=1e-4, betas=(0.9, 0.99), eps=1e-8, max_norm=1.0, ema_beta=0.995, ema_update_every=10,
                 ema_update_after_step=100, milestones=(50000, 150000, 300000), lr_gamma=0.1, group=None, bucket_mb=25):
        self.gd, self.model = diffusion, diffusion.model
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self.ema_beta, self.ema_every, self.ema_after = ema_beta, ema_update_every, ema_update_after_step
        self.milestones, self.lr_gamma = milestones, lr_gamma
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.step_count, self.ema_calls = 0, 0
        self.p = flat_params(self.model)
        self.g = flat_grads(self.model)
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.ema = self.p.clone()
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.p.device)
        self._params = [p for p in self.model.parameters() if p.requires_grad]
        self._off = dict(zip((id(p) for p in self._params), _offsets(self._params)))
        n = self.p.numel()
        per = max(1, int(bucket_mb * (1 << 20) / 4))
        self.buckets = [(lo, min(n, lo + per)) for lo in range(0, n, per)]
        self._pending = None

    # ------------------------------------------------------------------ gradient all-reduce, bucketed and overlapped
    def _bucket_plan(self, te):
        """bucket b is launched after the LAST tape record (in backward order) that finalises one of its parameters; buckets
        holding late parameters (time-embedding MLPs via autograd, the shared relative-position embedding) go after backward"""
        final_at = {}
        order = list(reversed(te.tape))
        for i, rec in enumerate(order):
            for p in te.record_params(rec):
                final_at[id(p)] = i
        rel = self.model.time_rel_pos_bias.relative_attention_bias.weight if hasattr(self.model, "time_rel_pos_bias") else None
        ready = {}
        for b, (lo, hi) in enumerate(self.buckets):
            last, late = -1, False
            for p in self._params:
                o = self._off[id(p)]
                if o < hi and o + p.numel() > lo:
                    if id(p) not in final_at or p is rel:
                        late = True
                    else:
                        last = max(last, final_at[id(p)])
            if not late:
                ready.setdefault(id(order[last]), []).append(b)
        return ready

    def _launch(self, b):
        lo, hi = self.buckets[b]
        self._works.append(dist.all_reduce(self.g[lo:hi], group=self.group, async_op=True))
        self._launched.add(b)

    def backward(self, loss):
        te = getattr(self.model, "_train_engine", None)
        self._works, self._launched = [], set()
        if self.world > 1 and te is not None and te.tape is not None:
            ready = self._bucket_plan(te)
            te.on_record_done = lambda rec: [self._launch(b) for b in ready.get(id(rec), ())]
        try:
            loss.backward()
        finally:
            if te is not None:
                te.on_record_done = None
        if self.world > 1:
            for b in range(len(self.buckets)):
                if b not in self._launched:
                    self._launch(b)
            for w in self._works:
                w.wait()
            self.g.mul_(1.0 / self.world)

    # ------------------------------------------------------------------ the step
    def current_lr(self):
        return self.lr * self.lr_gamma ** sum(self.step_count >= m for m in self.milestones)

    def step(self, state, *args, **kwargs):
        """loss = diffusion(state) -> backward -> all-reduce -> clip + Adam + EMA.  Returns the (detached) loss."""
        self.g.zero_()
        loss = self.gd(state, *args, **kwargs)
        self.backward(loss)
        self.optimizer_step()
        return loss.detach()

    def optimizer_step(self):
        L = _lib.lib()
        st = _lib.current_stream_ptr()
        self.sumsq.zero_()
        _lib.check(L.wdno_sumsq(_p(self.g), self.g.numel(), _p(self.sumsq), st), "sumsq")
        self.step_count += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.step_count, 1 - b2 ** self.step_count
        lr = self.lr * self.lr_gamma ** sum((self.step_count - 1) >= m for m in self.milestones)
        # ema_pytorch: update() is called once per optimiser step; acts every `update_every` calls; copies until update_after_step
        self.ema_calls += 1
        mode, decay = 0, 0.0
        if self.ema_calls % self.ema_every == 0:
            if self.ema_calls <= self.ema_after:
                mode = 1
            else:
                mode, decay = 2, ema_decay(self.ema_calls, self.ema_beta, self.ema_after)
        _lib.check(L.wdno_adam_clip_ema(_p(self.p), _p(self.g), _p(self.m), _p(self.v), _p(self.ema), self.p.numel(), _p(self.sumsq),
                                        float(self.max_norm), float(lr), float(b1), float(b2), float(self.eps), float(bc1),
                                        float(bc2), float(decay), mode, st), "adam_clip_ema")
        # the kernel wrote the parameters through raw pointers (no torch version bump): mark the engine's snapshot stale so
        # that model.engine() re-packs the tiles in place before the next forward
        if getattr(self.model, "_engine", None) is not None and self.model._engine_fp is not None:
            self.model._engine_fp = (-1, self.model._engine_fp[1])
        te = getattr(self.model, "_train_engine", None)
        if te is not None:
            te._versions = None

    def grad_norm(self):
        return float(self.sumsq.sqrt())

    def ema_state_dict(self):
        """EMA weights as a state dict of the model (non-trainable entries copied from the online model)"""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        named = dict(self.model.named_parameters())
        for k, p in named.items():
            if p.requires_grad:
                o = self._off[id(p)]
                sd[k] = self.ema[o:o + p.numel()].view_as(p).clone()
        return sd
