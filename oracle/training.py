"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of ONE training step of the reference:
the parity gate for the training kernels (not built in round 1; `p_losses` on the engine returns the forward value only).

    Trainer.train (smoke/ddpm/diffusion_2d.py:1257-1307 ; burgers/ddpm_burgers/train_diffusion.py:187-237):
        loss = model(state) / gradient_accumulate_every ; backward
        clip_grad_norm_(model.parameters(), 1.0)
        Adam(lr, betas = adam_betas).step() ; zero_grad ; (MultiStepLR [50000, 150000, 300000] x 0.1)
        ema.update()

* loss and gradients: autograd through the functional oracle networks (oracle/unet3d.py, oracle/unet2d.py), which are pinned
  bit-exactly to the reference modules; pinned here against `loss.backward()` of the real reference `GaussianDiffusion`
  (tests/test_oracle_training.py, build container only).
* clip / Adam: the documented torch algorithms (`torch.nn.utils.clip_grad_norm_`, `torch.optim.Adam` without weight decay /
  amsgrad), pinned against torch itself in the same test.
* EMA: `ema_pytorch.EMA(model, beta, update_every)` is NOT in the reference tree and not installable here (requirements.txt:
  unpinned) -> restated from its published algorithm with its defaults (update_after_step = 100, inv_gamma = 1, power = 2/3,
  min_value = 0): "parity unpinned" for the warm-up schedule.
"""
import torch

from . import diffusion as D
from .unet2d import Unet2DOracle
from .unet3d import Unet3DOracle


def trainable_names(state_dict):
    """parameters the reference optimises: everything in the U-Net state dict except the rotary frequency table
    (rotary-embedding-torch keeps `freqs` as a parameter with requires_grad = False)"""
    return [k for k in state_dict if not k.endswith("rotary_emb.freqs")]


def smoke_loss_and_grads(state_dict, sch, x_start, t, noise, coef_shape, loss_layer_weight, control=True, pad=True,
                         super_model=False, accumulate_every=1):
    """-> (loss, {name: dloss/dparam}) of smoke `GaussianDiffusion.forward` given the drawn (t, noise)
    (diffusion_2d.py:988-1058, 1278-1283)"""
    orc = Unet3DOracle(state_dict)
    names = trainable_names(orc.sd)
    for k in names:
        orc.sd[k] = orc.sd[k].clone().requires_grad_()
    loss = D.smoke_p_losses(orc, sch, x_start, t, noise, coef_shape, loss_layer_weight, control=control, pad=pad,
                            super_model=super_model) / accumulate_every
    grads = torch.autograd.grad(loss, [orc.sd[k] for k in names], allow_unused=True)
    return loss.detach(), {k: (torch.zeros_like(orc.sd[k]) if g is None else g) for k, g in zip(names, grads)}


def burgers_loss_and_grads(state_dict, sch, x_start, t, noise, coef_shape, loss_layer_weight, cond_u0=True, cond_uT=False,
                           cond_f=True, pad=True, super_model=False, accumulate_every=1):
    """-> (loss, {name: dloss/dparam}) of Burgers `GaussianDiffusion.forward` given the drawn (t, noise)
    (burgers/ddpm_burgers/diffusion_1d.py:520-654 ; train_diffusion.py:200-212).  Every entry of the Unet2D state dict is
    trainable (no rotary table)."""
    orc = Unet2DOracle(state_dict)
    names = list(orc.sd)
    for k in names:
        orc.sd[k] = orc.sd[k].clone().requires_grad_()
    loss = D.burgers_p_losses(orc, sch, x_start, t, noise, coef_shape, loss_layer_weight, cond_u0=cond_u0, cond_uT=cond_uT,
                              cond_f=cond_f, pad=pad, super_model=super_model) / accumulate_every
    grads = torch.autograd.grad(loss, [orc.sd[k] for k in names], allow_unused=True)
    return loss.detach(), {k: (torch.zeros_like(orc.sd[k]) if g is None else g) for k, g in zip(names, grads)}


def clip_grad_norm(grads, max_norm=1.0, eps=1e-6):
    """torch.nn.utils.clip_grad_norm_ (2-norm): -> (total_norm, clipped grads)"""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads.values()]), 2.0)
    coef = torch.clamp(max_norm / (total + eps), max=1.0)
    return total, {k: g * coef for k, g in grads.items()}


def adam_init(params):
    return {"step": 0, "m": {k: torch.zeros_like(v) for k, v in params.items()}, "v": {k: torch.zeros_like(v) for k, v in params.items()}}


def adam_update(params, grads, st, lr, betas=(0.9, 0.99), eps=1e-8):
    """torch.optim.Adam, default flags: m, v moments, bias corrections, p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)"""
    b1, b2 = betas
    st["step"] += 1
    bc1, bc2 = 1 - b1 ** st["step"], 1 - b2 ** st["step"]
    out = {}
    for k, p in params.items():
        g = grads[k]
        m = st["m"][k].mul_(b1).add_(g, alpha=1 - b1)
        v = st["v"][k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        out[k] = p - (lr / bc1) * (m / denom)
    return out


def multistep_lr(base_lr, step, milestones=(50000, 150000, 300000), gamma=0.1):
    """lr used by optimizer step number `step` (0-based) under MultiStepLR (diffusion_2d.py:1137-1138)"""
    return base_lr * gamma ** sum(step >= m for m in milestones)


def ema_decay(step, beta=0.995, update_after_step=100, inv_gamma=1.0, power=2.0 / 3.0, min_value=0.0):
    """ema_pytorch.EMA.get_current_decay (published algorithm; unpinned)"""
    epoch = max(step - update_after_step - 1, 0)
    if epoch <= 0:
        return 0.0
    return min(max(1 - (1 + epoch / inv_gamma) ** -power, min_value), beta)


def ema_update(ema, params, step, beta=0.995, update_every=10, update_after_step=100):
    """ema_pytorch.EMA.update, called once per optimizer step with its own counter `step` (0-based, before increment):
    every `update_every` calls; copies the online parameters until `update_after_step`, then lerps with the current decay"""
    step += 1
    if step % update_every != 0:
        return ema, step
    if step <= update_after_step:
        return {k: v.clone() for k, v in params.items()}, step
    d = ema_decay(step, beta, update_after_step)
    return {k: ema[k] + (1.0 - d) * (params[k] - ema[k]) for k in params}, step


def smoke_train_step(params, opt, sch, x_start, t, noise, coef_shape, loss_layer_weight, lr, betas=(0.9, 0.99), control=True,
                     pad=True, super_model=False, max_norm=1.0):
    """one optimizer step; params: full U-Net state dict (the rotary table rides along unchanged) -> (loss, total grad
    norm, new params)"""
    loss, grads = smoke_loss_and_grads(params, sch, x_start, t, noise, coef_shape, loss_layer_weight, control, pad, super_model)
    total, grads = clip_grad_norm(grads, max_norm)
    names = list(grads)
    new = adam_update({k: params[k] for k in names}, grads, opt, lr, betas)
    out = dict(params)
    out.update(new)
    return loss, total, out
