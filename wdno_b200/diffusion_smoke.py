"""smoke GaussianDiffusion on the B200 engine -- same constructor / sample() / p_losses() / forward() surface as
/root/reference/smoke/ddpm/diffusion_2d.py:568-1058 (class GaussianDiffusion).

The sampling loop keeps the state `x` [B,F,C,H,W] fp32 resident in HBM; one step is
    eps = Unet3D engine(x, t)   ->   fused (x0 clamp, eps re-derivation, DDIM/DDPM update, noise, conditions)
captured once in a CUDA graph and replayed per step (per-step scalars come from device tables).
"""
import math
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import ops


def linear_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    ac = torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1, clamp_min=1e-5):
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def register_schedule(module, betas):
    """the 13 fp32 buffers both reference GaussianDiffusion classes register (diffusion_2d.py:627-685)."""
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
    reg = lambda name, val: module.register_buffer(name, val.to(torch.float32))
    reg("betas", betas)
    reg("alphas_cumprod", ac)
    reg("alphas_cumprod_prev", ac_prev)
    reg("sqrt_alphas_cumprod", torch.sqrt(ac))
    reg("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - ac))
    reg("log_one_minus_alphas_cumprod", torch.log(1.0 - ac))
    reg("sqrt_recip_alphas_cumprod", torch.sqrt(1.0 / ac))
    reg("sqrt_recipm1_alphas_cumprod", torch.sqrt(1.0 / ac - 1))
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    reg("posterior_variance", pv)
    reg("posterior_log_variance_clipped", torch.log(pv.clamp(min=1e-20)))
    reg("posterior_mean_coef1", betas * torch.sqrt(ac_prev) / (1.0 - ac))
    reg("posterior_mean_coef2", (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac))
    return alphas, ac


_TABLE_CACHE = {}


def ddim_tables(mod, eta, gscale_fn=None):
    """per-step scalars of ddim_sample (diffusion_2d.py:862-911), computed with the same fp32 torch scalar ops.
    -> (times list, coef table [S, 8] fp32 cpu).  Cached per (schedule buffers, S, eta): ~10 scalar torch ops per step on
    the host would otherwise sit at the head of every sample() call."""
    T, S = mod.num_timesteps, mod.sampling_timesteps
    key = ("ddim", mod.alphas_cumprod.data_ptr(), mod.alphas_cumprod._version, T, S, float(eta))
    if key in _TABLE_CACHE:
        times, tab = _TABLE_CACHE[key]
        tab = tab.clone()
        if gscale_fn is not None:
            for i, t in enumerate(times):
                tab[i, 6] = gscale_fn(t)
        return list(times), tab
    times, tab = _ddim_tables(mod, eta, None)
    _TABLE_CACHE[key] = (list(times), tab.clone())
    if gscale_fn is not None:
        for i, t in enumerate(times):
            tab[i, 6] = gscale_fn(t)
    return times, tab


def _ddim_tables(mod, eta, gscale_fn=None):
    T, S = mod.num_timesteps, mod.sampling_timesteps
    times = torch.linspace(-1, T - 1, steps=S + 1)
    times = list(reversed(times.int().tolist()))
    pairs = list(zip(times[:-1], times[1:]))
    ac = mod.alphas_cumprod.detach().float().cpu()
    sr = mod.sqrt_recip_alphas_cumprod.detach().float().cpu()
    srm1 = mod.sqrt_recipm1_alphas_cumprod.detach().float().cpu()
    tab = torch.zeros(len(pairs), 8, dtype=torch.float32)
    for i, (t, tn) in enumerate(pairs):
        tab[i, 0], tab[i, 1] = sr[t], srm1[t]
        if tn < 0:
            tab[i, 5] = 1.0
        else:
            alpha, alpha_next = ac[t], ac[tn]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            tab[i, 2], tab[i, 3], tab[i, 4] = alpha_next.sqrt(), c, sigma
        if gscale_fn is not None:
            tab[i, 6] = gscale_fn(t)
    return [p[0] for p in pairs], tab


def ddpm_tables(mod, gscale_fn=None):
    """per-step scalars of p_sample (diffusion_2d.py:714-721,769-785) for t = T-1 .. 0."""
    T = mod.num_timesteps
    g = lambda name: getattr(mod, name).detach().float().cpu()
    sr, srm1 = g("sqrt_recip_alphas_cumprod"), g("sqrt_recipm1_alphas_cumprod")
    c1, c2, lv = g("posterior_mean_coef1"), g("posterior_mean_coef2"), g("posterior_log_variance_clipped")
    times = list(reversed(range(T)))
    tab = torch.zeros(T, 8, dtype=torch.float32)
    for i, t in enumerate(times):
        tab[i, 0], tab[i, 1], tab[i, 2], tab[i, 3] = sr[t], srm1[t], c1[t], c2[t]
        tab[i, 4] = (0.5 * lv[t]).exp() if t > 0 else 0.0
        if gscale_fn is not None:
            tab[i, 6] = gscale_fn(t)
    return times, tab


class StepRunner:
    """Runs `n` sampling steps: eps = model(x, t); fused update.  Un-guided steps are replayed from one CUDA graph."""

    def __init__(self, model, x, times, table, prog, kind, cond_mode, use_graph=True, capacity=None):
        self.model, self.x, self.kind, self.cond_mode, self.prog = model, x, kind, cond_mode, prog
        dev = x.device
        self.B = x.shape[0]
        # device tables with room for `capacity` steps: the captured graphs bake the table ADDRESSES and the capacity, so a
        # new schedule (another sampling_timesteps / eta) is an upload, not a re-capture
        self.cap = max(int(capacity or 0), len(times))
        self.time_table = torch.zeros(self.cap, dtype=torch.float32, device=dev)
        self.coef_table = torch.zeros(self.cap, 8, dtype=torch.float32, device=dev)
        self.set_schedule(times, table)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.time_f = torch.zeros(self.B, dtype=torch.float32, device=dev)
        self.coef = torch.zeros(8, dtype=torch.float32, device=dev)
        self.noise = torch.empty_like(x)
        self.graph = None
        self.use_graph = use_graph
        self.launches_per_step = None

    def set_schedule(self, times, table):
        n = len(times)
        assert n <= self.cap, "schedule longer than the runner's table capacity"
        self.times, self.n = list(times), n
        self.time_table[:n].copy_(torch.tensor(times, dtype=torch.float32), non_blocking=True)
        self.coef_table[:n].copy_(table.to(torch.float32), non_blocking=True)

    def _net(self):
        """eps = model(x, t) with the engine told that t is batch-uniform (step_begin writes one value to every row)"""
        eng = self.model.engine() if hasattr(self.model, "engine") else None
        if eng is None or not hasattr(eng, "forward") or os.environ.get("WDNO_TIME_UNIFORM", "1") == "0":
            return self.model(self.x, self.time_f)
        eng.time_uniform = True
        try:
            return self.model(self.x, self.time_f)
        finally:
            eng.time_uniform = False

    def _body(self, with_noise, guidance=None):
        ops.step_begin(self.step, self.time_table, self.coef_table, self.time_f, self.coef, self.cap)
        eps = self._net()
        fn = ops.ddim_step if self.kind == "ddim" else ops.ddpm_step
        fn(self.x, eps, self.noise if with_noise else None, self.coef, self.prog, self.cond_mode, guidance=guidance)
        return eps

    def step_eager(self, with_noise):
        self._body(with_noise)

    def guided_head(self):
        """step_begin -> U-Net -> x0 prediction of a guided step: -> (eps, x0).  Everything before the user's callback is
        replayed from a CUDA graph of its own (static eps / x0 buffers), so a guided step costs one graph launch + the
        callback + the fused update instead of ~125 eager launches."""
        clip = self.kind == "ddim"

        def head():
            ops.step_begin(self.step, self.time_table, self.coef_table, self.time_f, self.coef, self.cap)
            eps = self._net()
            return eps, ops.predict_x0(self.x, eps, self.coef, clip=clip)
        if not self.use_graph:
            return head()
        if getattr(self, "_head_graph", None) is None:
            s0 = self.step.clone()
            head()                      # warm-up outside capture (lazy plans, table uploads); x is not modified
            torch.cuda.synchronize()
            self.step.copy_(s0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._head_out = head()
            self._head_graph = g
        self._head_graph.replay()
        return self._head_out

    def _guided_body(self, with_noise, design, head_graph):
        """design(x0) -> gradient tensor (user code; torch autograd) ; eps += gscale * g inside the fused kernel."""
        if head_graph:
            eps, x0 = self.guided_head()
        else:
            ops.step_begin(self.step, self.time_table, self.coef_table, self.time_f, self.coef, self.cap)
            eps = self._net()
            x0 = ops.predict_x0(self.x, eps, self.coef, clip=self.kind == "ddim")
        g = design(x0)
        if not torch.is_tensor(g):
            g = None
        else:
            g = g.detach().to(torch.float32).contiguous()
        fn = ops.ddim_step if self.kind == "ddim" else ops.ddpm_step
        fn(self.x, eps, self.noise if with_noise else None, self.coef, self.prog, self.cond_mode, guidance=g)

    def step_guided(self, with_noise, design, graph=False):
        """One guided step.  Default: graph of the head + eager callback + fused update.  graph=True (opt-in,
        `GaussianDiffusion.graph_design_fn` / WDNO_GRAPH_GUIDANCE=1): the WHOLE step, callback included, is captured
        once per `design` closure and replayed -- valid when the callback is pure device work on its argument (the stock
        objectives of inference_2d.py / eval_ddpm_burgers.py are: transforms, reductions, autograd.grad; no host reads).
        A callback that cannot be captured (host synchronisation) falls back to the default path for good."""
        if graph and self.use_graph and not getattr(self, "_gg_failed", False):
            cache = getattr(self, "_gg", None)
            if cache is None or cache[0] is not design:   # the closure binds this call's init / low / init_u tensors
                cache = self._gg = (design, {})
            graphs = cache[1]
            if with_noise not in graphs:
                s0, xs = self.step.clone(), self.x.clone()
                try:
                    self._guided_body(with_noise, design, False)   # warm-up outside capture, then rewind
                    torch.cuda.synchronize()
                    self.step.copy_(s0)
                    self.x.copy_(xs)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._guided_body(with_noise, design, False)
                    graphs[with_noise] = g
                except Exception as e:  # noqa: BLE001 - user code under capture
                    import warnings
                    warnings.warn(f"guided step not capturable ({type(e).__name__}: {e}); running the callback eagerly")
                    self._gg_failed = True
                    torch.cuda.synchronize()
                    self.step.copy_(s0)
                    self.x.copy_(xs)
            if with_noise in graphs:
                graphs[with_noise].replay()
                return
        self._guided_body(with_noise, design, True)

    def step_graph(self, with_noise):
        if not self.use_graph:
            return self.step_eager(with_noise)
        if self.graph is None:
            self.graph = {}
        if with_noise not in self.graph:
            # warm-up outside capture (lazy plan building, table uploads), then rewind the step counter
            s0 = self.step.clone()
            x0 = self.x.clone()
            self._body(with_noise)
            torch.cuda.synchronize()
            self.step.copy_(s0)
            self.x.copy_(x0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body(with_noise)
            # capture does not execute; state untouched
            self.graph[with_noise] = g
        self.graph[with_noise].replay()


class GaussianDiffusion(nn.Module):
    def __init__(self, model, loss_layer_weight, is_condition_control, is_condition_pad, is_wavelet, is_super_model,
                 wave_type, pad_mode, padded_shape, ori_shape, *, image_size, frames, timesteps=1000,
                 sampling_timesteps=None, loss_type="l2", beta_schedule="sigmoid", schedule_fn_kwargs=dict(),
                 ddim_sampling_eta=0.0, min_snr_loss_weight=False, min_snr_gamma=5, standard_fixed_ratio=0.01,
                 coeff_ratio=0.1, objective=None):
        # `objective` is accepted and ignored: smoke/train_2d.py:120 passes it although the reference signature lacks it
        super().__init__()
        self.model = model
        self.loss_layer_weight = loss_layer_weight
        self.is_condition_control = is_condition_control
        self.is_condition_pad = is_condition_pad
        self.channels = self.model.channels
        self.self_condition = self.model.self_condition
        self.image_size = image_size
        self.frames = frames
        self.is_wavelet = is_wavelet
        self.is_super_model = is_super_model
        self.wave_type = wave_type
        self.pad_mode = pad_mode
        self.padded_shape = padded_shape
        self.ori_shape = ori_shape
        self.standard_fixed_ratio = standard_fixed_ratio
        self.coeff_ratio = coeff_ratio
        if beta_schedule == "linear":
            fn = linear_beta_schedule
        elif beta_schedule == "cosine":
            fn = cosine_beta_schedule
        elif beta_schedule == "sigmoid":
            fn = sigmoid_beta_schedule
        else:
            raise ValueError(f"unknown beta schedule {beta_schedule}")
        betas = fn(timesteps, **schedule_fn_kwargs)
        alphas, ac = register_schedule(self, betas)
        self.num_timesteps = int(betas.shape[0])
        self.loss_type = loss_type
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        snr = ac / (1 - ac)
        clipped = snr.clone()
        if min_snr_loss_weight:
            clipped.clamp_(max=min_snr_gamma)
        self.register_buffer("loss_weight", (clipped / snr).to(torch.float32))
        self.use_cuda_graph = True
        self.graph_design_fn = os.environ.get("WDNO_GRAPH_GUIDANCE", "0") == "1"   # capture design_fn with the step (opt-in)
        self._noise_source = None  # tests: callable(shape, device) replacing torch.randn (injected noise)
        self._step_hook = None     # tests: callable(step index, state) after every sampling step (trajectory traces)
        self.last_launches_per_step = None

    # ------------------------------------------------------------ helpers
    def _randn(self, shape, device):
        if self._noise_source is not None:
            return self._noise_source(tuple(shape), device).to(device=device, dtype=torch.float32).contiguous()
        return torch.randn(shape, device=device)

    def sample_noise(self, shape, device):
        return self._randn(shape, device)

    def _coef_shape(self, N_upsample):
        if not self.is_super_model:
            return self.padded_shape
        ps = self.padded_shape[N_upsample]
        if self.is_condition_control:
            return [ps[0], ps[1] + 2, ps[2] + 2]
        return [ps[0] + 2, ps[1], ps[2]]

    def _conditions(self, shape, coef_shape, init, control, low):
        """the in-place slice assignments of diffusion_2d.py:869-888 / 913-929 as an ordered condition program"""
        b, f, c, h, w = shape
        dev = self.betas.device
        prog = ops.CondProgram()
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        assert init is not None
        init = f32(init)
        if self.is_wavelet:
            assert init.shape == (b, f, h, w), f"init must be [B,F,H,W] = {(b, f, h, w)}, got {tuple(init.shape)}"
            prog.copy(init, "bfyx", c=(-2, -1))
        else:
            assert init.shape == (b, h, w)
            prog.copy(init, "byx", f=(0, 1), c=(0, 1))
        if self.is_condition_control:
            control = f32(control)
            if self.is_wavelet:
                assert control.shape == (b, f, 16, h, w)
                prog.copy(control, "bfcyx", c=(24, 40))
            else:
                assert control.shape == (b, f, 2, h, w)
                prog.copy(control, "bfcyx", c=(3, 5))
        if self.is_condition_pad and self.is_wavelet:
            T, Hh, Ww = coef_shape[-3], coef_shape[-2], coef_shape[-1]
            prog.zero(f=(T, None), c=(0, -2))
            prog.zero(f=(T, None), c=(-1, None))
            prog.zero(c=(0, -1), y=(Hh, None))
            prog.zero(c=(0, -1), x=(Ww, None))
        if self.is_super_model:
            low = f32(low)
            assert low.shape == (b, f, 40, h, w)
            prog.copy(low, "bfcyx", c=(40, 80))
        return prog, prog.build(f, c, h, w)

    def _guidance(self, design_fn, design_guidance, low, init, init_u):
        if design_fn is None:
            return None, None
        if design_guidance == "standard":
            gs = lambda t: self.standard_fixed_ratio
        elif design_guidance == "standard-alpha":
            sched = self.coeff_ratio * self.betas.detach().float().cpu().flip(0)
            gs = lambda t: float(sched[t])
        else:
            raise ValueError(f"unsupported design_guidance {design_guidance!r}")

        def design(x0):
            with torch.enable_grad():
                xc = x0.clone().detach().requires_grad_()
                return design_fn(xc, low=low, init=init, init_u=init_u)
        return design, gs

    # ------------------------------------------------------------ samplers
    def _runner(self, kind, shape, N_upsample, init, control, low, gs):
        """StepRunner with STATIC state / noise / condition buffers, cached across sample() calls so the captured
        CUDA graph is reused; a new call only copies its conditions into the static buffers."""
        dev = self.betas.device
        key = (kind, tuple(shape), N_upsample, control is not None and self.is_condition_control, low is not None,
               self.use_cuda_graph)
        sched_key = (kind, self.sampling_timesteps, float(self.ddim_sampling_eta))
        cache = self.__dict__.setdefault("_runners", {})
        r = cache.get(key)
        if r is not None and r.model_engine is self.model.engine() and r.sched_key != sched_key:
            # same workload, another schedule: upload the new per-step tables; the captured graphs stay valid
            times, table = ddim_tables(self, self.ddim_sampling_eta, gs) if kind == "ddim" else ddpm_tables(self, gs)
            r.set_schedule(times, table)
            r.sched_key = sched_key
        if r is None or r.model_engine is not self.model.engine():
            coef_shape = self._coef_shape(N_upsample)
            static = dict(init=torch.empty_like(init, dtype=torch.float32, device=dev).contiguous(),
                          control=None if control is None else torch.empty_like(control, dtype=torch.float32, device=dev).contiguous(),
                          low=None if low is None else torch.empty_like(low, dtype=torch.float32, device=dev).contiguous())
            keep, prog = self._conditions(shape, coef_shape, static["init"], static["control"], static["low"])
            if kind == "ddim":
                times, table = ddim_tables(self, self.ddim_sampling_eta, gs)
            else:
                times, table = ddpm_tables(self, gs)
            x = torch.empty(tuple(shape), dtype=torch.float32, device=dev)
            r = StepRunner(self.model, x, times, table, prog, kind, 1 if kind == "ddim" else 2,
                           use_graph=self.use_cuda_graph, capacity=self.num_timesteps)
            r.static, r.keep, r.model_engine, r.sched_key = static, keep, self.model.engine(), sched_key
            cache.clear()  # one live runner: its static buffers are sized for the workload
            cache[key] = r
        for k, v in (("init", init), ("control", control), ("low", low)):
            if r.static[k] is not None:
                r.static[k].copy_(v, non_blocking=True)
        if gs is not None:
            # the guidance scale depends on design_guidance / standard_fixed_ratio / coeff_ratio of THIS call (the reference
            # recomputes it every step, diffusion_2d.py:733-747): refresh column 6 of the device table the cached graphs read
            col = torch.tensor([float(gs(t)) for t in r.times], dtype=torch.float32)
            r.coef_table[:r.n, 6].copy_(col.to(dev), non_blocking=True)
        r.step.zero_()
        return r

    def _bound_design(self, run, design_fn, design_guidance, low, init, init_u):
        """the guidance callback bound to the runner's STATIC copies of (low, init, init_u): the closure -- and with it a
        whole-step CUDA graph that captured it -- survives across sample() calls; a call only refreshes the buffers"""
        if design_fn is None:
            return None
        dev = self.betas.device
        st = run.static
        if init_u is not None:
            if st.get("init_u") is None or st["init_u"].shape != init_u.shape:
                st["init_u"] = torch.empty(init_u.shape, dtype=torch.float32, device=dev)
                run._design_key = None
            st["init_u"].copy_(init_u, non_blocking=True)
        else:
            st["init_u"] = None
        lo = st["low"] if low is not None and st.get("low") is not None else low
        if lo is not None and lo is low:
            if st.get("low_g") is None or st["low_g"].shape != low.shape:
                st["low_g"] = torch.empty(low.shape, dtype=torch.float32, device=dev)
                run._design_key = None
            st["low_g"].copy_(low, non_blocking=True)
            lo = st["low_g"]
        key = (id(design_fn), design_guidance, lo is None)
        if getattr(run, "_design_key", None) != key:
            design, _ = self._guidance(design_fn, design_guidance, lo, st["init"], st["init_u"])
            run._design, run._design_key, run._design_fn = design, key, design_fn
            run._gg = None
        return run._design

    @torch.no_grad()
    def ddim_sample(self, shape, N_upsample=0, design_fn=None, design_guidance="standard", init=None, init_u=None,
                    control=None, low=None, device=None):
        dev = self.betas.device
        _, gs = self._guidance(design_fn, design_guidance, low, init, init_u)
        assert init is not None
        run = self._runner("ddim", shape, N_upsample, init, control if self.is_condition_control else None,
                           low if self.is_super_model else None, gs)
        design = self._bound_design(run, design_fn, design_guidance, low, init, init_u)
        run.x.copy_(self._randn(shape, dev))
        ops.apply_conditions(run.x, run.prog)
        n = len(run.times)
        for i in range(n):
            last = i == n - 1
            if not last:
                run.noise.copy_(self._randn(shape, dev)) if self._noise_source is not None else run.noise.normal_()
            if design is not None:
                run.step_guided(not last, design, graph=self.graph_design_fn)
            else:
                run.step_graph(not last)
            if self._step_hook is not None:
                self._step_hook(i, run.x)
        self.last_launches_per_step = self.model.engine().launches + 2
        return run.x.clone()

    @torch.no_grad()
    def p_sample_loop(self, shape, N_upsample=0, design_fn=None, design_guidance="standard", return_all_timesteps=None,
                      init=None, init_u=None, control=None, low=None, device=None):
        dev = self.betas.device
        _, gs = self._guidance(design_fn, design_guidance, low, init, init_u)
        assert init is not None
        run = self._runner("ddpm", shape, N_upsample, init, control if self.is_condition_control else None,
                           low if self.is_super_model else None, gs)
        design = self._bound_design(run, design_fn, design_guidance, low, init, init_u)
        run.x.copy_(self._randn(list(shape), dev))
        ops.apply_conditions(run.x, run.prog)
        for i, t in enumerate(run.times):
            with_noise = t > 0
            if with_noise:
                run.noise.copy_(self._randn(shape, dev)) if self._noise_source is not None else run.noise.normal_()
            if design is not None:
                run.step_guided(with_noise, design, graph=self.graph_design_fn)
            else:
                run.step_graph(with_noise)
            if self._step_hook is not None:
                self._step_hook(i, run.x)
        self.last_launches_per_step = self.model.engine().launches + 2
        return run.x.clone()

    @torch.no_grad()
    def sample(self, batch_size=16, N_upsample=0, design_fn=None, design_guidance="standard", init=None, init_u=None,
               control=None, low=None, device=None):
        assert batch_size == init.shape[0]
        if not self.is_super_model:
            size = (batch_size, self.frames, self.channels, self.image_size, self.image_size)
        else:
            size = (batch_size, low.shape[1], self.channels, low.shape[-2], low.shape[-1])
        fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return fn(size, N_upsample, design_fn, design_guidance, init=init, init_u=init_u, control=control, low=low,
                  device=device)

    # ------------------------------------------------------------ training objective (forward value)
    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = self._randn(x_start.shape, x_start.device)
        return ops.q_sample(x_start.contiguous().float(), noise.contiguous().float(), self.sqrt_alphas_cumprod,
                            self.sqrt_one_minus_alphas_cumprod, t.contiguous())

    def p_losses(self, state_start, t, noise=None):
        """diffusion_2d.py:988-1050.  With autograd enabled and trainable parameters the returned loss carries a grad_fn whose
        backward runs on the engine (train3d.py); under torch.no_grad() it is the forward value only."""
        b, f, c, h, w = state_start.shape
        if self.is_super_model:
            if self.is_condition_control:
                nd = int(math.log2(40 / w))
                ps = self.padded_shape[nd]
                coef_shape = [ps[0], ps[1] + 2, ps[2] + 2]
            else:
                nd = int(math.log2(24 / f))
                ps = self.padded_shape[nd]
                coef_shape = [ps[0] + 2, ps[1], ps[2]]
        else:
            coef_shape = self.padded_shape
        state_start = state_start.contiguous().float()
        noise_state = noise if noise is not None else self._randn(state_start.shape, state_start.device)
        noise_state = noise_state.contiguous().float()
        state = self.q_sample(state_start, t, noise_state)
        # conditioned slices <- clean data, their noise target <- 0; padded region <- 0 in both
        ps_, pn_ = ops.CondProgram(), ops.CondProgram()
        zeros_src = None
        if self.is_wavelet:
            ps_.copy(state_start, "bfcyx", c=(-2, -1))
            pn_.zero(c=(-2, -1))
        else:
            ps_.copy(state_start, "bfcyx", f=(0, 1), c=(0, 1))
            pn_.zero(f=(0, 1), c=(0, 1))
        if self.is_condition_control:
            cr = (24, 40) if self.is_wavelet else (3, 5)
            ps_.copy(state_start, "bfcyx", c=cr)
            pn_.zero(c=cr)
        if self.is_condition_pad and self.is_wavelet:
            T, Hh, Ww = coef_shape[-3], coef_shape[-2], coef_shape[-1]
            for pr in (ps_, pn_):
                pr.zero(f=(T, None), c=(0, -2))
                pr.zero(f=(T, None), c=(-1, None))
                pr.zero(c=(0, -1), y=(Hh, None))
                pr.zero(c=(0, -1), x=(Ww, None))
        if self.is_super_model:
            ps_.copy(state_start, "bfcyx", c=(40, 80))
            pn_.zero(c=(40, 80))
        # CondProgram.copy offsets the source by the box origin; for "same tensor" copies shift the pointer back
        ops.apply_conditions(state, _same_tensor_program(ps_, state_start, f, c, h, w))
        ops.apply_conditions(noise_state, pn_.build(f, c, h, w))
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()):
            # training: differentiable forward, backward in libwdno_b200.so (wdno_b200/train3d.py); `loss.backward()` fills
            # p.grad exactly like the reference's autograd step (diffusion_2d.py:1277-1284)
            if self.loss_type != "l2":
                raise NotImplementedError("l1 loss is never used by WDNO")
            from .train3d import unet3d_apply
            model_out = unet3d_apply(self.model, state, t)
            loss = F.mse_loss(model_out, noise_state, reduction="mean")
            lw = self.loss_layer_weight
            if torch.is_tensor(lw):
                lw = lw.to(loss.device)
            loss = loss * lw
            return loss.mean() if torch.is_tensor(loss) else loss
        with torch.no_grad():
            model_out = self.model(state, t)
        if self.loss_type == "l2":
            acc = ops.mse_weighted(model_out, noise_state, None)
            loss = (acc.sum() / model_out.numel()).to(torch.float32)
        elif self.loss_type == "l1":
            raise NotImplementedError("l1 loss is never used by WDNO")
        else:
            raise ValueError(f"invalid loss type {self.loss_type}")
        lw = self.loss_layer_weight
        if torch.is_tensor(lw):
            lw = lw.to(loss.device)
        loss = loss * lw
        return loss.mean() if torch.is_tensor(loss) else loss

    def forward(self, state, *args, **kwargs):
        b = state.shape[0]
        t = torch.randint(0, self.num_timesteps, (b,), device=state.device).long()
        return self.p_losses(state, t, *args, **kwargs)


def _same_tensor_program(prog, src, f, c, h, w):
    """build a program whose COPY ops read `src` at the SAME index as the destination (src has the state's shape)."""
    arr, n = prog.build(f, c, h, w)
    for i in range(n):
        o = arr[i]
        if o.src:
            o.of, o.oc, o.oy, o.ox = 0, 0, 0, 0
    return arr, n
