"""Sampling glue of the Burgers experiment on the B200 engine: guidance and the zero-shot
super-resolution cascade (row f-2).  The reference's versions are tied to its dataset files (`get_target`) and its
numerical solver; here the same arithmetic takes the target / condition tensors as arguments:

    ddpm_guidance_loss(u_target, u, f, wu, wf, condition_f)     burgers/ddpm_burgers/test_util.py:100-126
    get_nablaJ(loss_fn)                                         burgers/ddpm_burgers/model_utils.py:35-50
    get_scheduler(name) + the four J schedules                  burgers/ddpm_burgers/model_utils.py:52-131
    get_loss_fn_2dconv(u_target, args, shape, ori_shape, ...)   burgers/eval_ddpm_burgers.py:108-143
    diffuse_fields(ddpm, args, RESCALER, **sample_kwargs)       burgers/eval_ddpm_burgers.py:151-193  (sample -> u, f)
    next_level_low(sampled_coef, padded_shape, k, is_wavelet)   burgers/eval_ddpm_burgers.py:305-309
    run_cascade(...)                                            burgers/eval_ddpm_burgers.py:279-338  (without metrics)

The U-Net / sampler run through GaussianDiffusion.sample(); the inverse transform and its autograd adjoint are
libwdno_b200.so kernels (wdno_b200.wavelets).  Solver metrics, result files and checkpoint loading are out of scope.
"""
import math

import torch
import torch.nn.functional as F

from wdno_b200.packing import burgers_tensor_to_coef as tensor_to_coef
from wdno_b200.packing import burgers_tensor_to_coef_super as tensor_to_coef_super
from wdno_b200.packing import burgers_upsample_coef as upsample_coef
from wdno_b200.wavelets import DWTInverse


# ------------------------------------------------------------------ guidance objective and its gradient
def ddpm_guidance_loss(u_target, u, f, wu=0, wf=0, condition_f=False):
    """u_target, u [B,Nt,Nx]; f [B,Nt-1,Nx] -> scalar (summed over the batch)"""
    d0 = (u[:, 0, :] - u_target[:, 0, :]).square()
    if not condition_f:
        d0 = d0 + (u[:, -1, :] - u_target[:, -1, :]).square()
    return (d0.mean(-1).sum() + f.square().sum() * wf) * wu


def get_nablaJ(loss_fn):
    """x -> dJ/dx (detached); J = loss_fn(x)"""
    def nablaJ(x):
        x.requires_grad_(True)
        J = loss_fn(x)
        g = torch.autograd.grad(J, x, grad_outputs=torch.ones_like(J), allow_unused=True)[0]
        return g.detach()
    return nablaJ


def _betas_cosine(s=0.008, T=1000):
    x = torch.linspace(0, T, T + 1, dtype=torch.float64)
    ac = torch.cos(((x / T) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def _betas_sigmoid(start=-3, end=3, tau=1, T=1000):
    x = torch.linspace(0, T, T + 1, dtype=torch.float64) / T
    v0, v1 = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
    ac = (-((x * (end - start) + start) / tau).sigmoid() + v1) / (v1 - v0)
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def cosine_beta_J_schedule(t, s=0.008):
    return _betas_cosine(s)[t]


def sigmoid_schedule(t, start=-3, end=3, tau=1, clamp_min=1e-5):
    return _betas_sigmoid(start, end, tau)[t]


def sigmoid_schedule_flip(t):
    return sigmoid_schedule(999 - t)


def get_scheduler(scheduler):
    if scheduler is None:
        return None
    table = {"cosine": cosine_beta_J_schedule, "sigmoid": sigmoid_schedule, "sigmoid_flip": sigmoid_schedule_flip}
    if scheduler == "linear":
        raise NotImplementedError
    if scheduler == "plain_cosine":
        raise NotImplementedError("the reference's plain_cosine_schedule calls Tensor.flip() without dims and raises")
    if scheduler not in table:
        raise ValueError(scheduler)
    return table[scheduler]


def get_loss_fn_2dconv(u_target, args, shape, ori_shape, RESCALER, is_super_model=False, wf=0, wu=0, condition_f=False):
    """u_target [B,Nt,Nx] physical target trajectory of this resolution level (the reference loads it from its test
    set); returns loss_fn(x) on the rescaled state x [B,C,H,W]"""
    def loss_fn(x):
        x = x[:, :8] * RESCALER[:, :8] if is_super_model and args.is_wavelet else x * RESCALER
        if not args.is_wavelet:
            return ddpm_guidance_loss(u_target[:, :shape[-2], :shape[-1]], x[:, 0, :shape[-2], :shape[-1]],
                                      x[:, 1, :shape[-2] - 1, :shape[-1]], wu=wu, wf=wf, condition_f=condition_f)
        Yl, Yh = tensor_to_coef(x, shape)
        u_f = DWTInverse(mode=args.pad_mode, wave=args.wave_type)((Yl, Yh))[:, :, :ori_shape[-2], :ori_shape[-1]]
        return ddpm_guidance_loss(u_target[:, :ori_shape[-2], :ori_shape[-1]], u_f[:, 0], u_f[:, 1, :ori_shape[-2] - 1],
                                  wu=wu, wf=wf, condition_f=condition_f)
    return loss_fn


def get_nablaJ_2dconv(**kwargs):
    return get_nablaJ(get_loss_fn_2dconv(**kwargs))


# ------------------------------------------------------------------ sampling + inverse transform
def diffuse_fields(ddpm, args, RESCALER=1, **kwargs):
    """-> (coefficients x[:, :8] (or x[:, :2]) in physical units cropped to the level's padded shape, u [B,Nt,Nx],
    f [B,Nt-1,Nx])"""
    if "N_upsample" not in kwargs:
        shape, ori_shape = ddpm.padded_shape, ddpm.ori_shape
    else:
        shape, ori_shape = ddpm.padded_shape[kwargs["N_upsample"] - 1], ddpm.ori_shape[kwargs["N_upsample"] - 1]
    x = ddpm.sample(**kwargs) * RESCALER
    if not ddpm.is_wavelet:
        return x[:, :2], x[:, 0, :shape[-2], :shape[-1]], x[:, 1, :shape[-2] - 1, :shape[-1]]
    Yl, Yh = (tensor_to_coef_super if "low" in kwargs else tensor_to_coef)(x, shape)
    x = x[:, :, :shape[-2], :shape[-1]]
    u_f = DWTInverse(mode=args.pad_mode, wave=args.wave_type)((Yl, Yh))[:, :, :ori_shape[-2], :ori_shape[-1]]
    return x[:, :8], u_f[:, 0], u_f[:, 1, :ori_shape[-2] - 1]


def coef_state_to_trajectory(x, shape, ori_shape, wave_type, pad_mode, RESCALER=1):
    """rescaled base-model state [B,9,64,64] (what `sample()` returns) -> physical fields [B,2,Nt,Nx] (u, f) through the
    inverse 2-D transform (eval_ddpm_burgers.py:188-194): the `post` of `wdno_b200.parallel.sample_sharded` for Burgers"""
    Yl, Yh = tensor_to_coef(x * RESCALER, shape)
    return DWTInverse(mode=pad_mode, wave=wave_type)((Yl, Yh))[:, :, :ori_shape[-2], :ori_shape[-1]]


def next_level_low(sampled_coef, padded_shape_k, k, is_wavelet=True):
    """nearest x2 of the previous level's coefficients, zero-padded to the level-k padded plane"""
    low = upsample_coef(sampled_coef, padded_shape_k)
    size = (64 if is_wavelet else 128) * 2 ** k
    return F.pad(low, (0, size - low.shape[-1], 0, size - low.shape[-2]), "constant", 0)


def run_cascade(ddpm, ddpm_super, args, RESCALER, u_targets, u_conditions, fs, wu=0, wf=0, J_scheduler=None):
    """base level + `args.upsample_x` zero-shot super-resolution levels.
    u_targets[k] [B,Nt_k,Nx_k]: physical target of level k; u_conditions[k]: wavelet-domain u condition rows
    (physical units); fs[k]: wavelet-domain f condition (physical units).  -> list of (coef, u, f) per level."""
    R_base = RESCALER[:, 8:17] if (args.is_super_model and args.is_wavelet) else RESCALER
    B = u_targets[0].shape[0]
    sched = get_scheduler(J_scheduler)
    uc = u_conditions[0]
    out = [diffuse_fields(
        ddpm, args, RESCALER=R_base, batch_size=B, J_scheduler=sched,
        u_init=uc[:, :32] / R_base.squeeze()[-1], u_final=uc[:, -32:] / R_base.squeeze()[-1], f=fs[0] / R_base[:, 4:8],
        nablaJ=get_nablaJ_2dconv(u_target=u_targets[0], args=args, shape=ddpm.padded_shape, ori_shape=ddpm.ori_shape,
                                 RESCALER=R_base, wu=wu, wf=wf, condition_f=args.is_condition_f))]
    if not args.is_super_model:
        return out
    coef = out[0][0]
    for k in range(1, args.upsample_x + 1):
        low = next_level_low(coef, ddpm_super.padded_shape[k - 1], k, args.is_wavelet) / RESCALER[:, 8:16]
        uc = u_conditions[k]
        res = diffuse_fields(
            ddpm_super, args, N_upsample=k, RESCALER=RESCALER, batch_size=B, J_scheduler=sched, low=low,
            u_init=uc[:, :32 * 2 ** k] / RESCALER.squeeze()[-1], u_final=uc[:, -32 * 2 ** k:] / RESCALER.squeeze()[-1],
            f=fs[k] / RESCALER[:, 4:8],
            nablaJ=get_nablaJ_2dconv(u_target=u_targets[k], args=args, shape=ddpm_super.padded_shape[k - 1],
                                     ori_shape=ddpm_super.ori_shape[k - 1], RESCALER=RESCALER, is_super_model=True,
                                     wu=0, wf=0, condition_f=args.is_condition_f))
        out.append(res)
        coef = res[0]
    return out
