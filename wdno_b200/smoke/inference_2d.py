"""Sampling glue of the smoke experiment on the B200 engine: guided sampling and the
super-resolution cascade (row f-2).  Same entry points and argument meaning as the reference's
/root/reference/smoke/inference_2d.py:
    guidance_fn(x, args, shape, ori_shape, RESCALER, w_energy, w_init, low, init, init_u)          (lines 30-66)
    make_design_fn(args, shape, ori_shape, RESCALER)            = the closure built in load_model   (lines 69-96)
    InferencePipeline.run_model / run_base_model / run_super_model                                  (lines 122-258)
Checkpoint loading (Trainer), the PhiFlow solver evaluation and result files are out of scope.

Everything numeric runs in libwdno_b200.so: the U-Net / DDIM loop through GaussianDiffusion.sample(), the wavelet
transforms through wdno_b200.wavelets (whose autograd backward is the adjoint kernel, so `torch.autograd.grad` of
the guidance objective never leaves the GPU and never differentiates through the U-Net: x0 is detached first,
diffusion_2d.py:733-747).
"""
import math

import torch
import torch.nn.functional as F

from wdno_b200.packing import smoke_coef_to_tensor as coef_to_tensor
from wdno_b200.packing import smoke_tensor_to_coef as tensor_to_coef
from wdno_b200.packing import smoke_upsample_coef as upsample_coef
from wdno_b200.wavelets import DWT1DInverse, DWTForward, Wavelet, wavedec3, waverec3, waverec3_adjoint

PAD_T, PAD_X = 24, 40          # padded frames / base padded width of the coefficient state


def _fields_from_state(wave_state, shape, ori_shape, wave_type, upsample_type=None):
    """rescaled-back coefficient state [B,F,>=40,H,W] -> physical fields [B,5,T,H,W] (rho, v1, v2, c1, c2)"""
    coef = tensor_to_coef(wave_state[:, :, :40].permute(0, 2, 1, 3, 4), shape, upsample_type=upsample_type)
    rec = waverec3(coef, Wavelet(wave_type))[:, :ori_shape[0], :ori_shape[1], :ori_shape[2]]
    return rec.reshape(-1, 5, ori_shape[0], ori_shape[1], ori_shape[2])


def _smoke_out_from_state(last_channel, n_coef, half, pad_mode, wave_type):
    """the smoke-out channel stores a 1-D transform: rows [:half] carry the low band, rows [half:] the high band, each
    replicated over the plane; the inverse gives the fraction of smoke leaving through the target per time step.
    last_channel [B,F,H,W] -> [B, T]"""
    lo = last_channel[:, :n_coef, :half].mean((-2, -1)).unsqueeze(1)
    hi = last_channel[:, :n_coef, half:].mean((-2, -1)).unsqueeze(1)
    return DWT1DInverse(mode=pad_mode, wave=wave_type)((lo, [hi]))[:, 0]


def state_to_fields(wave_output, RESCALER, shape, ori_shape, wave_type, pad_mode, upsample_type=None):
    """rescaled coefficient state [B,F,C,H,W] (what `sample()` returns) -> physical output [B,T,6,H,W]: 5 fields through
    the inverse 3-D transform + the smoke-out series through the inverse 1-D transform, broadcast over the plane
    (inference_2d.py:136-152).  This is the `post` of `wdno_b200.parallel.sample_sharded`: the fields are what is gathered."""
    scaled40 = wave_output[:, :, :40] * RESCALER[:, :, :40]
    fields = _fields_from_state(scaled40, shape, ori_shape, wave_type, upsample_type).permute(0, 2, 1, 3, 4)
    last = wave_output[:, :, -1] * RESCALER[:, :, -1]
    so = _smoke_out_from_state(last, shape[0], int(last.shape[-2] / 2), pad_mode, wave_type)
    so = so.reshape(so.shape[0], so.shape[1], 1, 1, 1).expand(-1, -1, -1, ori_shape[1], ori_shape[2])
    return torch.cat((fields, so), dim=2)


def guidance_fn(x, args, shape, ori_shape, RESCALER, w_energy=0, w_init=0, low=None, init=None, init_u=None):
    """gradient of the design objective J; `low`, `init` are rescaled, `init_u` is not (reference docstring).
    The reference rebinds `x = x * RESCALER` before differentiating (inference_2d.py:35,65), so what it returns is
    dJ/d(x * RESCALER) -- the gradient in PHYSICAL coefficient units, without the chain factor RESCALER -- and that is
    what the sampler adds to eps.  Reproduced as is."""
    xs = x * RESCALER
    if args.is_wavelet:
        fields = _fields_from_state(xs[:, :, :-2], shape, ori_shape, args.wave_type)
        smoke_out = _smoke_out_from_state(xs[:, :, -1], shape[0], 20, args.pad_mode, args.wave_type)  # rows split at 40/2 (line 44)
        j_success = smoke_out[:, ori_shape[0] - 1].sum()
        j_init = (fields[:, 0, 0] - init_u.to(fields.device)).square().mean((-1, -2)).sum()
        j_energy = fields[:, 3:5].square().mean((1, 2, 3, 4)).sum()
        if args.is_condition_control:
            j = w_init * j_init
        else:
            j = -j_success + w_energy * j_energy + w_init * j_init
    else:
        if args.is_condition_control:
            return torch.zeros_like(x)  # the reference differentiates the constant 0 here (which raises in autograd)
        j = -xs[:, -1, -1].mean((-1, -2)).sum() + w_energy * xs[:, 3:5].square().mean((1, 2, 3, 4)).sum()
    return torch.autograd.grad(j, xs, grad_outputs=torch.ones_like(j))[0]


_SMOKE_OUT_GRAD = {}


def _smoke_out_grad(n_coef, n_out, pad_mode, wave_type, device):
    """d smoke_out[n_out - 1] / d(lo, hi): the 1-D inverse transform is linear, so this is a constant pair of [n_coef] vectors
    (one autograd pass through DWT1DInverse, cached)"""
    key = (n_coef, n_out, pad_mode, wave_type, str(device))
    if key not in _SMOKE_OUT_GRAD:
        with torch.enable_grad():
            lo = torch.zeros(1, 1, n_coef, device=device, requires_grad=True)
            hi = torch.zeros(1, 1, n_coef, device=device, requires_grad=True)
            y = DWT1DInverse(mode=pad_mode, wave=wave_type)((lo, [hi]))[:, 0]
            g_lo, g_hi = torch.autograd.grad(y[:, n_out - 1].sum(), (lo, hi))
        _SMOKE_OUT_GRAD[key] = (g_lo.reshape(-1).detach(), g_hi.reshape(-1).detach())
    return _SMOKE_OUT_GRAD[key]


def guidance_fn_closed_form(x, args, shape, ori_shape, RESCALER, w_energy=0, w_init=0, low=None, init=None, init_u=None):
    """The same gradient as `guidance_fn` WITHOUT autograd (the objective is quadratic / linear in
    the reconstructed fields, so dJ/dfields is known in closed form and dJ/d(x * RESCALER) is its image under the adjoint
    inverse transform): one `waverec3`, a handful of elementwise writes, one `waverec3_adjoint`.
        J = -smoke_out[T - 1] + w_energy mean(c^2) + w_init mean((rho(t = 0) - rho0*)^2)      (summed over the batch)
        dJ/dfields[:, 3:5] = 2 w_energy c / (2 T H W) ;  dJ/dfields[:, 0, 0] = 2 w_init (rho0 - rho0*) / (H W)
        dJ/d(last channel): the constant pair d smoke_out[T - 1] / d(lo, hi) spread over the rows each mean was taken over
    Wavelet states only; anything else falls back to `guidance_fn`."""
    if not args.is_wavelet:
        return guidance_fn(x, args, shape, ori_shape, RESCALER, w_energy, w_init, low, init, init_u)
    with torch.no_grad():
        xs = x.detach() * RESCALER
        B = xs.shape[0]
        T, H, Wd = int(shape[-3]), int(shape[-2]), int(shape[-1])
        o0, o1, o2 = int(ori_shape[0]), int(ori_shape[1]), int(ori_shape[2])
        coef = tensor_to_coef(xs[:, :, :40].permute(0, 2, 1, 3, 4), shape)
        rec = waverec3(coef, Wavelet(args.wave_type))
        fields = rec[:, :o0, :o1, :o2].reshape(B, 5, o0, o1, o2)
        g_rec = torch.zeros_like(rec).reshape(B, 5, *rec.shape[1:])
        g_rec[:, 0, 0, :o1, :o2] = (2.0 * w_init / (o1 * o2)) * (fields[:, 0, 0] - init_u.to(fields.device))
        if not args.is_condition_control and w_energy != 0:
            g_rec[:, 3:5, :o0, :o1, :o2] = (2.0 * w_energy / (2 * o0 * o1 * o2)) * fields[:, 3:5]
        g_bands = waverec3_adjoint(g_rec.reshape(-1, *rec.shape[1:]), Wavelet(args.wave_type), (T, H, Wd))
        g40 = coef_to_tensor(g_bands).reshape(B, 40, T, H, Wd).permute(0, 2, 1, 3, 4)   # [B, T, 40, H, W]
        g = torch.zeros_like(xs)
        g[:, :T, :40, :H, :H] = g40[..., :H]   # the reference crops rows and columns with the same bound (tensor_to_coef)
        if not args.is_condition_control:
            g_lo, g_hi = _smoke_out_grad(T, o0, args.pad_mode, args.wave_type, xs.device)
            Hx, Wx = xs.shape[-2], xs.shape[-1]
            g[:, :T, -1, :20, :] = (-g_lo / (20 * Wx)).reshape(1, T, 1, 1)
            g[:, :T, -1, 20:, :] = (-g_hi / ((Hx - 20) * Wx)).reshape(1, T, 1, 1)
    return g


def make_design_fn(args, shape, ori_shape, RESCALER, closed_form=None):
    """the `design_fn(x, low=, init=, init_u=)` closure of inference_2d.py:82-91.  The gradient is evaluated with
    `guidance_fn_closed_form` (one waverec3, elementwise field gradient, one waverec3_adjoint: measured 5.60 ms per C5 step vs
    5.92 ms through autograd, profiles/r2_bench_configs.jsonl; same values, tests/test_gpu_pipeline.py); closed_form=False or
    WDNO_CLOSED_FORM_GUIDANCE=0 selects `guidance_fn` (torch.autograd.grad of the reference's objective through the kernels)."""
    if closed_form is None:
        import os
        closed_form = os.environ.get("WDNO_CLOSED_FORM_GUIDANCE", "1") != "0"
    guidance = guidance_fn_closed_form if closed_form else guidance_fn

    def design_fn(x, low=None, init=None, init_u=None):
        kw = dict(w_energy=args.w_energy, w_init=args.w_init, low=low, init=init, init_u=init_u)
        if args.is_super_model:
            if low is not None:
                lvl = int(math.log2(low.shape[-1] / 40))
                return guidance(x, args, shape[lvl], ori_shape[lvl], RESCALER, **kw)
            return guidance(x, args, shape[0], ori_shape[0], RESCALER[:, :, 40:], **kw)
        return guidance(x, args, shape, ori_shape, RESCALER, **kw)
    return design_fn


class InferencePipeline(object):
    """model: [base diffusion] or [base diffusion, super diffusion]; args: dict(design_fn=, design_guidance=);
    args_general: namespace with is_wavelet, is_condition_control, image_size, device, upsample, is_super_model."""

    def __init__(self, model, args=None, RESCALER=1, results_path=None, args_general=None):
        self.model = model
        self.args = args
        self.results_path = results_path
        self.args_general = args_general
        self.is_wavelet = args_general.is_wavelet
        self.is_condition_control = args_general.is_condition_control
        self.image_size = args_general.image_size
        self.device = args_general.device
        self.upsample = args_general.upsample
        self.RESCALER = RESCALER

    # ------------------------------------------------------------ wavelet conditions of one resolution level
    def _wave_init(self, rho0, gd, nx, pad_x, level):
        """2-D transform of the initial density -> (LL, LH) repeated over the padded frames (base: 4 blocks of 6
        frames; super levels: interleaved), zero-padded to the padded plane"""
        yl, yh = DWTForward(J=1, mode=gd.pad_mode, wave=gd.wave_type)(rho0)
        w = torch.cat((yl, yh[0][:, 0]), dim=1)  # [B,4,nx,nx]: LL and the three detail bands of channel 0
        if level == 0:
            w = w.unsqueeze(2).expand(w.shape[0], 4, PAD_T // 4, nx, nx)
        else:
            w = w.unsqueeze(1).expand(w.shape[0], PAD_T // 4, 4, nx, nx)
        w = w.reshape(-1, PAD_T, nx, nx)
        return F.pad(w, (0, pad_x - nx, 0, pad_x - nx), "constant", 0)

    def _wave_control(self, control, gd, pad_x, level, n_t):
        """3-D transform of the two control fields -> 16 channels [B,F,16,H,W] on the padded grid; super levels are
        replicate-padded by one coefficient (space) or one frame (time) first"""
        b, t = control.shape[0], control.shape[1]
        vol = control.permute(0, 2, 1, 3, 4).reshape(-1, t, control.shape[3], control.shape[4])
        w = coef_to_tensor(wavedec3(vol, Wavelet(gd.wave_type), mode=gd.pad_mode, level=1))
        if level > 0:
            if not self.args_general.is_condition_control:
                w = torch.cat((w[:, :, :1], w[:, :, :n_t], w[:, :, [n_t - 1]]), dim=2)
            else:
                w = F.pad(w.reshape(-1, *w.shape[-3:]), (1, 1, 1, 1), mode="replicate")
        w = F.pad(w, (0, pad_x - w.shape[-1], 0, pad_x - w.shape[-2], 0, PAD_T - w.shape[-3]), "constant", 0)
        return w.reshape(-1, 2, 8, PAD_T, pad_x, pad_x).reshape(-1, 16, PAD_T, pad_x, pad_x).permute(0, 2, 1, 3, 4)

    def _to_fields(self, wave_output, shape, ori_shape, gd, upsample_type=None):
        """rescaled state [B,F,C,H,W] -> physical output [B,T,6,H,W] (5 fields + smoke-out broadcast over the plane)"""
        return state_to_fields(wave_output, self.RESCALER, shape, ori_shape, gd.wave_type, gd.pad_mode, upsample_type)

    # ------------------------------------------------------------ samplers
    def _sample(self, gd, state, wave_init, wave_control, **kw):
        return gd.sample(batch_size=state.shape[0], design_fn=self.args["design_fn"],
                         design_guidance=self.args["design_guidance"],
                         init=wave_init / self.RESCALER[:, :, -2], init_u=state[:, 0, 0],
                         control=wave_control / self.RESCALER[:, :, 24:40], **kw)

    def run_base_model(self, state, wave_init, wave_control):
        gd = self.model[0]
        if not self.is_wavelet:
            out = gd.sample(batch_size=state.shape[0], design_fn=self.args["design_fn"],
                            design_guidance=self.args["design_guidance"], low=None,
                            init=state[:, 0, 0] / self.RESCALER[:, 0, 0], init_u=state[:, 0, 0],
                            control=state[:, :, 3:5] / self.RESCALER[:, :, 3:5]) * self.RESCALER
            out[:, :, -1] = out[:, :, -1].mean((-2, -1)).unsqueeze(-1).unsqueeze(-1).expand(-1, -1, 64, 64)
            return out
        out = self._sample(gd, state, wave_init, wave_control, low=None)
        return self._to_fields(out, gd.padded_shape, gd.ori_shape, gd)

    def run_super_model(self, state, state_ori, wave_init, wave_control):
        base, sup = self.model[0], self.model[1]
        shape, ori_shape = sup.padded_shape, sup.ori_shape
        assert self.args_general.is_condition_control
        up = "space"
        unpack = lambda w, sh, ut=None: coef_to_tensor(tensor_to_coef(w[:, :, :40].permute(0, 2, 1, 3, 4), sh, upsample_type=ut)) \
            .reshape(-1, 5, 8, *sh).reshape(-1, 40, *sh).permute(0, 2, 1, 3, 4)
        outs = [self._sample(base, state, wave_init, wave_control, low=None)]
        rets = [unpack(outs[0], shape[0])]
        for i in range(1, self.upsample + 1):
            pad_x, nx = PAD_X * 2 ** i, sup.padded_shape[i][-2]
            if not self.args_general.is_condition_control:
                state = state_ori[:, ::2 ** (3 - i)]
            else:
                s = 2 ** (1 - i)
                state = state_ori[:, :, :, ::s, ::s] if s >= 1 else state_ori
            w_init = self._wave_init(state[:, 0, [0]], sup, nx, pad_x, i)
            w_ctrl = self._wave_control(state[:, :, 3:5], sup, pad_x, i, shape[i][0])
            low = upsample_coef(rets[-1], shape[i], type=up)
            low = F.pad(low, (0, pad_x - low.shape[-1], 0, pad_x - low.shape[-2], 0, 0, 0, PAD_T - low.shape[-4]),
                        "constant", 0)
            outs.append(self._sample(sup, state, w_init, w_ctrl, N_upsample=i, low=low))
            rets.append(unpack(outs[-1], shape[i], up))
        return [self._to_fields(o, shape[i], ori_shape[i], base, None if i == 0 else up) for i, o in enumerate(outs)]

    def run_model(self, state):
        """state: physical fields, not rescaled, [B, T, 6, H, W] at the finest resolution of the cascade"""
        state_ori = state.to(self.device)
        state = state_ori[:, ::8] if not self.args_general.is_condition_control else state_ori[:, :, :, ::2, ::2]
        wave_init = wave_control = None
        if self.is_wavelet:
            gd = self.model[0]
            shp = gd.padded_shape
            wave_init = self._wave_init(state[:, 0, [0]], gd, shp[-2], PAD_X, 0)
            wave_control = self._wave_control(state[:, :, 3:5], gd, PAD_X, 0, shp[0])
        if len(self.model) == 1:
            return self.run_base_model(state, wave_init, wave_control)
        return self.run_super_model(state, state_ori, wave_init, wave_control)
