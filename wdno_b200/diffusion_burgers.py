"""Burgers GaussianDiffusion / GaussianDiffusion1D on the B200 engine -- same constructor kwargs, sample(**kwargs),
p_losses() and forward() as /root/reference/burgers/ddpm_burgers/diffusion_1d.py:40-658.

State layout [B, C, nt, nx] fp32 (treated as [B, 1, C, nt, nx] by the fused step kernels).  The reference applies
`set_condition` (pad, u0, uT, f, low) at the START of every iteration and once more after the loop
(diffusion_1d.py:395-415, 437-457); that is the same as imposing them on the initial noise and after every update,
which is what the fused DDIM/DDPM step kernel does (cond_mode = 2).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .diffusion_smoke import StepRunner, cosine_beta_schedule, ddim_tables, ddpm_tables, linear_beta_schedule, register_schedule


def _default_ori_shape():
    # the reference default is torch.tensor([81, 128], device='cuda') (diffusion_1d.py:51); it is never read on the hot path
    return torch.tensor([81, 128])


class GaussianDiffusion(nn.Module):
    def __init__(self, model, *, seq_length, is_wavelet=True, pad_mode=None, wave_type=None, padded_shape=None,
                 ori_shape=None, is_super_model=False, upsample_t=1, upsample_x=1, timesteps=1000,
                 sampling_timesteps=None, objective="pred_noise", beta_schedule="cosine", ddim_sampling_eta=0.0,
                 auto_normalize=False, loss_layer_weight=1, is_condition_pad=True, is_condition_u0=False,
                 is_condition_uT=False, is_condition_f=False, train_on_padded_locations=True):
        super().__init__()
        self.is_wavelet, self.is_super_model = is_wavelet, is_super_model
        self.pad_mode, self.wave_type = pad_mode, wave_type
        self.model = model
        self.channels = self.model.channels
        self.self_condition = self.model.self_condition
        self.traj_size = seq_length
        self.objective = objective
        assert objective in {"pred_noise", "pred_x0", "pred_v"}, \
            "objective must be either pred_noise (predict noise) or pred_x0 (predict image start) or pred_v"
        if objective != "pred_noise":
            raise NotImplementedError("every WDNO script trains/samples with objective='pred_noise'")
        if auto_normalize:
            raise NotImplementedError("auto_normalize is False in every WDNO script")
        if beta_schedule == "linear":
            betas = linear_beta_schedule(timesteps)
        elif beta_schedule == "cosine":
            betas = cosine_beta_schedule(timesteps)
        else:
            raise ValueError(f"unknown beta schedule {beta_schedule}")
        alphas, ac = register_schedule(self, betas)
        self.alphas = alphas.to(torch.float32).clone()
        self.alphas_prev = torch.nn.functional.pad(alphas[:-1], (1, 0), value=1.0).to(torch.float32).clone()
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        snr = ac / (1 - ac)
        self.register_buffer("loss_weight", torch.ones_like(snr).to(torch.float32))
        self.normalize = self.unnormalize = lambda t, *a, **k: t
        self.loss_layer_weight = loss_layer_weight
        self.upsample_t, self.upsample_x = upsample_t, upsample_x
        self.is_condition_pad = is_condition_pad
        self.is_condition_u0, self.is_condition_uT, self.is_condition_f = is_condition_u0, is_condition_uT, is_condition_f
        self.train_on_padded_locations = train_on_padded_locations
        self.padded_shape = padded_shape
        self.ori_shape = ori_shape if ori_shape is not None else _default_ori_shape()
        self.use_cuda_graph = True
        self._noise_source = None
        self._step_hook = None     # tests: callable(step index, state) after every sampling step
        self.last_launches_per_step = None

    # ------------------------------------------------------------ helpers
    def _randn(self, shape, device):
        if self._noise_source is not None:
            return self._noise_source(tuple(shape), device).to(device=device, dtype=torch.float32).contiguous()
        return torch.randn(shape, device=device)

    def get_guidance_options(self, **kwargs):
        nabla_J = kwargs.get("nablaJ")
        if nabla_J is not None:
            assert not self.self_condition, "self condition not tested with guidance"
        sched = kwargs.get("J_scheduler") or (lambda t: 1.0)
        proj = kwargs.get("proj_guidance") or (lambda ep, nj: ep + nj)
        return nabla_J, sched, proj

    def _coef_shape(self, kwargs):
        if not self.is_super_model:
            return self.padded_shape
        ps = self.padded_shape[kwargs["N_upsample"] - 1]
        return [ps[0] + 1, ps[1]]

    def _program(self, shape, coef_shape, srcs):
        """set_condition sequence pad, u0, uT, f, low (diffusion_1d.py:276-307) as an ordered condition program.
        srcs: dict of STATIC fp32 CUDA tensors (u0 / uT / f / low) or None."""
        b, c, H, W = shape
        Hh, Ww = int(coef_shape[-2]), int(coef_shape[-1])
        prog = ops.CondProgram()
        if self.is_wavelet:
            if self.is_condition_pad:
                prog.zero(c=(0, -1), y=(Hh, None))
                prog.zero(x=(Ww, None))
            if srcs.get("u0") is not None:
                u = srcs["u0"]
                prog.copy(u, "byx", c=(-1, None), y=(0, u.shape[-2]), x=(0, Ww))
            if srcs.get("uT") is not None:
                u = srcs["uT"]
                prog.copy(u, "byx", c=(-1, None), y=(H - u.shape[-2], H), x=(0, Ww))
            if srcs.get("f") is not None:
                prog.copy(srcs["f"], "bcyx", c=(4, 8), y=(0, Hh), x=(0, Ww))
            if srcs.get("low") is not None:
                prog.copy(srcs["low"], "bcyx", c=(8, 16), y=(0, Hh), x=(0, Ww))
        else:
            if self.is_condition_pad:
                prog.zero(c=(0, 1), y=(Hh, None))
                prog.zero(c=(1, 2), y=(Hh - 1, None))
                prog.zero(x=(Ww, None))
            if srcs.get("u0") is not None:
                prog.copy(srcs["u0"], "bx", c=(0, 1), y=(0, 1), x=(0, Ww))
            if srcs.get("uT") is not None:
                u = srcs["uT"]
                if u.dim() == 3:
                    prog.copy(u, "byx", c=(0, 1), y=(Hh - 2, Hh), x=(0, Ww))
                else:
                    prog.copy(u, "bx", c=(0, 1), y=(Hh - 1, Hh), x=(0, Ww))
            if srcs.get("f") is not None:
                prog.copy(srcs["f"], "byx", c=(1, 2), y=(0, Hh - 1), x=(0, Ww))
            if srcs.get("low") is not None:
                prog.copy(srcs["low"], "bcyx", c=(2, 4), y=(0, Hh), x=(0, Ww))
        return prog, prog.build(1, c, H, W)

    def _gather_sources(self, kwargs, dev):
        f32 = lambda t: t.to(device=dev, dtype=torch.float32)
        src = {}
        if self.is_condition_u0:
            u = f32(kwargs["u_init"])
            src["u0"] = u if self.is_wavelet else u.reshape(u.shape[0], -1)
        if self.is_condition_uT:
            u = f32(kwargs["u_final"])
            if not self.is_wavelet:
                u = u.unsqueeze(1).expand(-1, 2, -1).contiguous() if self.is_super_model else u.reshape(u.shape[0], -1)
            src["uT"] = u
        if self.is_condition_f:
            f = f32(kwargs["f"])
            src["f"] = f if self.is_wavelet else f.reshape(f.shape[0], f.shape[-2], f.shape[-1])
        if self.is_super_model:
            src["low"] = f32(kwargs["low"])
        return src

    def _runner(self, kind, shape, kwargs):
        dev = self.betas.device
        srcs = self._gather_sources(kwargs, dev)
        key = (kind, tuple(shape), kwargs.get("N_upsample"), tuple((k, tuple(v.shape)) for k, v in srcs.items()),
               self.use_cuda_graph)
        sched_key = (kind, self.sampling_timesteps, float(self.ddim_sampling_eta))
        cache = self.__dict__.setdefault("_runners", {})
        r = cache.get(key)
        if r is not None and r.model_engine is self.model.engine() and r.sched_key != sched_key:
            times, table = (ddim_tables(self, self.ddim_sampling_eta) if kind == "ddim" else ddpm_tables(self))
            r.set_schedule(times, table)
            r.sched_key = sched_key
        if r is None or r.model_engine is not self.model.engine():
            static = {k: torch.empty(v.shape, dtype=torch.float32, device=dev) for k, v in srcs.items()}
            keep, prog = self._program(shape, self._coef_shape(kwargs), static)
            times, table = (ddim_tables(self, self.ddim_sampling_eta) if kind == "ddim" else ddpm_tables(self))
            x = torch.empty((shape[0], 1) + tuple(shape[1:]), dtype=torch.float32, device=dev)
            r = StepRunner(_As5D(self.model), x, times, table, prog, kind, 2, use_graph=self.use_cuda_graph,
                           capacity=self.num_timesteps)
            r.static, r.keep, r.model_engine, r.sched_key = static, keep, self.model.engine(), sched_key
            cache.clear()
            cache[key] = r
        for k, v in srcs.items():
            r.static[k].copy_(v, non_blocking=True)
        r.step.zero_()
        return r

    def _guided_step(self, run, with_noise, nabla_J, sched, proj):
        """eps' = proj_guidance(eps, nablaJ(x0) * J_scheduler(t))  (user callables, torch), then the fused update"""
        eps, x0 = run.guided_head()   # step_begin -> U-Net -> x0, replayed from its own CUDA graph
        t = int(run.times[int(run.step.item()) - 1])
        with torch.enable_grad():
            g = nabla_J(x0[:, 0]) * sched(t)
            eps2 = proj(eps[:, 0], g)
        eps2 = eps2.detach().to(torch.float32).reshape(eps.shape).contiguous()
        fn = ops.ddim_step if run.kind == "ddim" else ops.ddpm_step
        fn(run.x, eps2, run.noise if with_noise else None, run.coef, run.prog, run.cond_mode)

    def _loop(self, kind, shape, kwargs):
        dev = self.betas.device
        nabla_J, sched, proj = self.get_guidance_options(**kwargs)
        run = self._runner(kind, shape, kwargs)
        shape5 = tuple(run.x.shape)
        run.x.copy_(self._randn(shape, dev).reshape(shape5))
        ops.apply_conditions(run.x, run.prog)
        n = len(run.times)
        for i, t in enumerate(run.times):
            with_noise = (i != n - 1) if kind == "ddim" else (t > 0)
            if with_noise:
                if self._noise_source is not None:
                    run.noise.copy_(self._randn(shape, dev).reshape(shape5))
                else:
                    run.noise.normal_()
            if nabla_J is not None:
                self._guided_step(run, with_noise, nabla_J, sched, proj)
            else:
                run.step_graph(with_noise)
            if self._step_hook is not None:
                self._step_hook(i, run.x.reshape(shape))
        self.last_launches_per_step = self.model.engine().launches + 2
        return run.x.reshape(shape).clone()

    @torch.no_grad()
    def p_sample_loop(self, shape, **kwargs):
        return self._loop("ddpm", shape, kwargs)

    @torch.no_grad()
    def ddim_sample(self, shape, **kwargs):
        return self._loop("ddim", shape, kwargs)

    def sample(self, batch_size=16, **kwargs):
        if self.is_condition_u0:
            assert "is_condition_u0" not in kwargs, "specify this value in the model. not during sampling."
            assert "u_init" in kwargs and kwargs["u_init"] is not None
        if self.is_condition_uT:
            assert "is_condition_uT" not in kwargs, "specify this value in the model. not during sampling."
            assert "u_final" in kwargs and kwargs["u_final"] is not None
        if self.is_condition_f:
            assert "is_condition_f" not in kwargs, "specify this value in the model. not during sampling."
            assert "f" in kwargs and kwargs["f"] is not None
        if self.is_super_model:
            assert "N_upsample" in kwargs and kwargs["N_upsample"] is not None
            assert "low" in kwargs and kwargs["low"] is not None
            size = (batch_size, self.channels, *kwargs["low"].shape[-2:])
        else:
            size = (batch_size, self.channels, *self.traj_size)
        fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return fn(size, **kwargs)

    # ------------------------------------------------------------ training objective (forward value)
    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = self._randn(x_start.shape, x_start.device)
        return ops.q_sample(x_start.contiguous().float(), noise.contiguous().float(), self.sqrt_alphas_cumprod,
                            self.sqrt_one_minus_alphas_cumprod, t.contiguous())

    def p_losses(self, x_start, t, noise=None):
        """Loss VALUE of diffusion_1d.py:529-645 through the engine (forward only; backward kernels are the 'next' row
        f-3).  Like the reference, the passed `noise` tensor is modified in place (conditioned targets are zeroed)."""
        b, c, nt, nx = x_start.shape
        if self.is_super_model:
            nd = int(math.log2((64 if self.is_wavelet else 128) / nx))
            coef_shape = [self.padded_shape[nd][0] + 1, self.padded_shape[nd][1]]
        else:
            coef_shape = self.padded_shape
        x_start = x_start.contiguous().float()
        if noise is None:
            noise = self._randn(x_start.shape, x_start.device)
        assert noise.is_contiguous() and noise.dtype == torch.float32
        x = self.q_sample(x_start, t, noise)
        half = int(nt / 2)
        if not self.is_wavelet:
            raise NotImplementedError("p_losses is built for the wavelet models (is_wavelet=True), the only ones WDNO trains")
        src = {}
        if self.is_condition_u0:
            src["u0"] = x_start[:, -1, :half, :].contiguous()
        if self.is_condition_uT:
            src["uT"] = x_start[:, -1, half:, :].contiguous()
        if self.is_condition_f:
            src["f"] = x_start[:, 4:8].contiguous()
        if self.is_super_model:
            src["low"] = x_start[:, 8:16].contiguous()
        _, prog_x = self._program((b, c, nt, nx), coef_shape, src)
        ops.apply_conditions(x.reshape(b, 1, c, nt, nx), prog_x)
        zsrc = {k: torch.zeros_like(v) for k, v in src.items()}
        _, prog_n = self._program((b, c, nt, nx), coef_shape, zsrc)
        ops.apply_conditions(noise.reshape(b, 1, c, nt, nx), prog_n)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()):
            # training: differentiable forward, backward in libwdno_b200.so (wdno_b200/train2d.py); diffusion_1d.py:639-645
            from .train2d import unet2d_apply
            out = unet2d_apply(self.model, x, t)
            loss = F.mse_loss(out, noise, reduction="none")
            lw = self.loss_layer_weight
            loss = loss * (lw.to(loss.device) if torch.is_tensor(lw) else lw)
            loss = loss.reshape(b, -1).mean(dim=1)
            return (loss * self.loss_weight[t]).mean()
        with torch.no_grad():
            out = self.model(x, t)
        lw = self.loss_layer_weight
        w = None
        if torch.is_tensor(lw):
            w = lw.to(device=out.device, dtype=torch.float32).reshape(-1).contiguous()
            assert w.numel() in (1, c), "loss_layer_weight must be a scalar or one weight per channel"
        acc = ops.mse_weighted(out, noise, w)
        per_sample = (acc / (c * nt * nx)).to(torch.float32)
        if not torch.is_tensor(lw):
            per_sample = per_sample * float(lw)
        return (per_sample * self.loss_weight[t]).mean()

    def forward(self, img, *args, **kwargs):
        b = img.shape[0]
        t = torch.randint(0, self.num_timesteps, (b,), device=img.device).long()
        return self.p_losses(img, t, *args, **kwargs)


class GaussianDiffusion1D(GaussianDiffusion):
    pass


class _As5D:
    """adapts Unet2D (x [B,C,H,W]) to the step runner's [B,1,C,H,W] state"""

    def __init__(self, model):
        self.model = model

    def __call__(self, x5, t):
        b, one, c, h, w = x5.shape
        return self.model(x5.reshape(b, c, h, w), t).reshape(b, 1, -1, h, w)

    def engine(self):
        return self.model.engine()
