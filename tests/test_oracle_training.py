"""Pins oracle/training.py (one training step: loss, gradients, gradient clipping, Adam, EMA schedule -- SURVEY.md section 8
row f-3, the parity gate of the training kernels) against the real reference `GaussianDiffusion` + torch's own
clip_grad_norm_ / Adam in the build container, and against the committed golden of the full-size model everywhere."""
import os

import pytest
import torch

from oracle import ref_loader
from oracle import training as TR

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smoke_train_step.pt")
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@needs_ref
def test_loss_gradients_clip_and_adam_match_reference_two_steps():
    from oracle import diffusion as D
    from tests.test_oracle_vs_reference import _smoke_pair
    s, m, gd, _ = _smoke_pair(S=None, eta=0.0, T=1000)
    gd.loss_layer_weight = torch.linspace(0.5, 2.0, 42).reshape(1, 1, 42, 1, 1)
    m.train()
    g = torch.Generator().manual_seed(4)
    x0 = torch.randn(2, 24, 42, 40, 40, generator=g).clamp(-1, 1)
    t = torch.tensor([17, 640])
    noise = torch.randn(2, 24, 42, 40, 40, generator=g)
    sch = D.schedule("sigmoid", 1000)
    lr = 1e-3
    opt_ref = torch.optim.Adam(gd.parameters(), lr=lr, betas=(0.9, 0.99))
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    opt = TR.adam_init({k: params[k] for k in TR.trainable_names(params)})
    for step in range(2):
        loss_ref = gd.p_losses(x0.clone(), t, noise.clone())
        loss_ref.backward()
        ref_grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
        loss, grads = TR.smoke_loss_and_grads(params, sch, x0, t, noise, [18, 34, 34], gd.loss_layer_weight)
        assert abs(float(loss) - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref)))
        assert set(ref_grads) == set(grads)   # same trainable set (the rotary table is not optimised)
        worst = max(rel_l2(grads[k], ref_grads[k]) for k in grads if float(ref_grads[k].norm()) > 1e-6)
        assert worst < 2e-4, worst
        total_ref = torch.nn.utils.clip_grad_norm_(gd.parameters(), 1.0)
        total, clipped = TR.clip_grad_norm(ref_grads, 1.0)
        assert abs(float(total) - float(total_ref)) < 1e-5 * float(total_ref)
        for n, p in m.named_parameters():
            if p.grad is not None:
                assert torch.allclose(clipped[n], p.grad, atol=1e-9, rtol=1e-6)
        opt_ref.step()
        opt_ref.zero_grad()
        # the Adam restatement fed with the reference's clipped gradients reproduces torch's update
        new = TR.adam_update({k: params[k] for k in clipped}, clipped, opt, lr, (0.9, 0.99))
        after = dict(m.named_parameters())
        for k, v in new.items():
            assert torch.allclose(v, after[k].detach(), atol=2e-7, rtol=1e-5), k
        params.update({k: after[k].detach().clone() for k in new})   # continue from the reference state


@needs_ref
def test_burgers_loss_and_gradients_match_reference():
    from oracle import diffusion as D
    from tests.test_oracle_vs_reference import _burgers_pair
    b, m, gd, _ = _burgers_pair(None, 0.0, T=1000)
    m.train()
    g = torch.Generator().manual_seed(6)
    x0 = torch.randn(2, 9, 64, 64, generator=g).clamp(-1, 1)
    t = torch.tensor([5, 911])
    noise = torch.randn(2, 9, 64, 64, generator=g)
    loss_ref = gd.p_losses(x0.clone(), t, noise.clone())
    loss_ref.backward()
    ref_grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    loss, grads = TR.burgers_loss_and_grads(params, D.schedule("cosine", 1000), x0, t, noise, [41, 60], gd.loss_layer_weight)
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref)))
    assert set(ref_grads) <= set(grads)
    worst = max(rel_l2(grads[k], ref_grads[k]) for k in ref_grads if float(ref_grads[k].norm()) > 1e-6)
    assert worst < 2e-4, worst
    for k in set(grads) - set(ref_grads):   # state-dict entries the reference does not optimise receive no gradient here either
        assert float(grads[k].abs().max()) == 0.0, k
    total_ref = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    total, _ = TR.clip_grad_norm(grads, 1.0)
    assert abs(float(total) - float(total_ref)) < 1e-4 * float(total_ref)


def test_ema_schedule_and_lr_milestones():
    assert TR.ema_decay(50) == 0.0 and TR.ema_decay(101) == 0.0
    assert abs(TR.ema_decay(110) - (1 - 10 ** (-2 / 3))) < 1e-12
    assert TR.ema_decay(10 ** 7) == 0.995
    p = {"w": torch.ones(3)}
    ema, n = {"w": torch.zeros(3)}, 0
    for _ in range(100):
        ema, n = TR.ema_update(ema, p, n)
    assert n == 100 and torch.equal(ema["w"], p["w"])     # copies the online weights during warm-up
    p2 = {"w": torch.full((3,), 3.0)}
    for _ in range(10):
        ema, n = TR.ema_update(ema, p2, n)
    d = TR.ema_decay(110)
    assert torch.allclose(ema["w"], torch.full((3,), 1.0 + (1 - d) * 2.0))
    assert TR.multistep_lr(1e-3, 0) == 1e-3 and abs(TR.multistep_lr(1e-3, 50000) - 1e-4) < 1e-12
    assert abs(TR.multistep_lr(1e-3, 300000) - 1e-6) < 1e-15


def test_oracle_train_step_reproduces_reference_golden():
    """full-size smoke model (dim 64), batch 1: the golden holds loss, clipped-gradient norms and strided gradient samples of
    the REAL reference backward (tests/golden/make_golden.py train); weights are re-created from the seed"""
    from oracle import diffusion as D
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    gold = torch.load(GOLD, weights_only=False)
    torch.manual_seed(0)
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert abs(float(sum(v.double().abs().sum() for v in sd.values())) - gold["weights_checksum"]) < 1e-6 * gold["weights_checksum"]
    g = torch.Generator().manual_seed(gold["input_seed"])
    x0 = torch.randn(1, 24, 42, 40, 40, generator=g).clamp(-1, 1)
    noise = torch.randn(1, 24, 42, 40, 40, generator=g)
    t = torch.tensor([gold["t"]])
    w = torch.linspace(0.5, 2.0, 42).reshape(1, 1, 42, 1, 1)
    loss, grads = TR.smoke_loss_and_grads(sd, D.schedule("sigmoid", 1000), x0, t, noise, [18, 34, 34], w)
    assert abs(float(loss) - gold["loss"]) < 1e-5 * abs(gold["loss"])
    total, _ = TR.clip_grad_norm(grads, 1.0)
    assert abs(float(total) - gold["total_norm"]) < 1e-4 * gold["total_norm"]
    for k, n in gold["grad_norms"].items():
        assert abs(float(grads[k].norm()) - n) <= 2e-4 * n + 1e-7, k
    for k, sub in gold["grad_subs"].items():
        assert rel_l2(grads[k].reshape(-1)[::gold["stride"]], sub) < 2e-4, k
