"""Host logic of the sampling glue (wdno_b200/smoke/inference_2d.py) on CPU: with the transforms swapped for the torch
oracle (oracle/wavelets_torch.py) and the REAL reference GaussianDiffusion / Unet3D plugged in as the sampler, our
InferencePipeline + guidance_fn must reproduce the goldens written by the reference's own InferencePipeline
(tests/golden/make_golden.py guided / cascade) to fp32 round-off -- i.e. the glue itself (condition preparation,
coefficient up-sampling, packing, the gradient convention) is the reference's, independent of any kernel."""
import os
import types

import pytest
import torch

from oracle import ref_loader
from oracle import wavelets_torch as T
from tests.test_oracle_vs_reference import NoiseTape, patched_randn

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture()
def inf(monkeypatch):
    from wdno_b200.smoke import inference_2d as m
    for n in ("waverec3", "wavedec3", "DWT1DInverse", "DWTForward", "Wavelet"):
        monkeypatch.setattr(m, n, getattr(T, n))

    def waverec3_adjoint(gy, wavelet, coef_shape):
        """adjoint of the (linear) oracle waverec3 by one autograd pass at zero"""
        leaves = [torch.zeros(gy.shape[0], *coef_shape, requires_grad=True) for _ in range(8)]
        with torch.enable_grad():
            y = T.waverec3([leaves[0], dict(zip(T.KEYS3, leaves[1:]))], wavelet)
            gs = torch.autograd.grad(y, leaves, gy)
        return [gs[0], dict(zip(T.KEYS3, gs[1:]))]
    monkeypatch.setattr(m, "waverec3_adjoint", waverec3_adjoint)
    monkeypatch.setattr(m, "_SMOKE_OUT_GRAD", {})
    return m


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _args(control, super_model):
    return types.SimpleNamespace(is_wavelet=True, wave_type="bior1.3", pad_mode="zero", is_condition_control=control,
                                 is_condition_pad=True, is_super_model=super_model, upsample=1 if super_model else 0,
                                 image_size=64, device="cpu", w_energy=0.5, w_init=0.1)


def test_guidance_gradient_and_base_pipeline_match_reference(inf):
    s = ref_loader.smoke()
    gold = torch.load(os.path.join(GOLD, "smoke_guided_pipeline.pt"))
    torch.manual_seed(0)
    m = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).eval()
    shape, ori_shape = [18, 34, 34], [32, 64, 64]
    rescaler = torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1)
    gd = s.GaussianDiffusion(m, rescaler, False, True, True, False, "bior1.3", "zero", shape, ori_shape, image_size=40,
                             frames=24, timesteps=1000, sampling_timesteps=gold["steps"], ddim_sampling_eta=1.0,
                             standard_fixed_ratio=100.0)
    args = _args(False, False)
    gen = torch.Generator().manual_seed(gold["input_seed"])
    state = 0.5 * torch.randn(1, 256, 6, 64, 64, generator=gen)
    xg = torch.randn(1, 24, 42, 40, 40, generator=gen).clamp(-1, 1).requires_grad_()
    g = inf.guidance_fn(xg, args, shape, ori_shape, rescaler, w_energy=0.5, w_init=0.1, init_u=state[:, 0, 0])
    assert rel_l2(g.reshape(-1)[::gold["stride"]], gold["grad_sub"]) < 1e-6
    pipe = inf.InferencePipeline([gd], args=dict(design_fn=inf.make_design_fn(args, shape, ori_shape, rescaler),
                                                 design_guidance="standard"), RESCALER=rescaler, args_general=args)
    with patched_randn(NoiseTape(gold["tape_seed"])), torch.no_grad():
        out = pipe.run_model(state)
    assert tuple(out.shape) == gold["out_shape"]
    assert rel_l2(out.reshape(-1)[::gold["stride"]], gold["out_sub"]) < 1e-5


def test_closed_form_guidance_gradient_equals_autograd_and_reference(inf):
    """guidance_fn_closed_form (no autograd: waverec3 -> closed-form field gradient -> adjoint transform) against our autograd
    guidance_fn and against the gradient the REAL reference guidance_fn produced (golden), control and design modes"""
    gold = torch.load(os.path.join(GOLD, "smoke_guided_pipeline.pt"))
    shape, ori_shape = [18, 34, 34], [32, 64, 64]
    rescaler = torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1)
    gen = torch.Generator().manual_seed(gold["input_seed"])
    state = 0.5 * torch.randn(1, 256, 6, 64, 64, generator=gen)
    xg = torch.randn(1, 24, 42, 40, 40, generator=gen).clamp(-1, 1)
    args = _args(False, False)
    g = inf.guidance_fn_closed_form(xg, args, shape, ori_shape, rescaler, w_energy=0.5, w_init=0.1, init_u=state[:, 0, 0])
    assert not g.requires_grad and rel_l2(g.reshape(-1)[::gold["stride"]], gold["grad_sub"]) < 1e-6
    assert abs(float(g.norm()) - gold["grad_norm"]) < 1e-5 * gold["grad_norm"]
    x2 = torch.randn(2, 24, 42, 40, 40, generator=gen).clamp(-1, 1)
    u2 = torch.randn(2, 64, 64, generator=gen)
    for control, we, wi in ((False, 0.5, 0.1), (False, 0.0, 0.3), (True, 0.7, 0.2)):
        a = _args(control, False)
        want = inf.guidance_fn(x2.clone().requires_grad_(), a, shape, ori_shape, rescaler, w_energy=we, w_init=wi, init_u=u2)
        got = inf.guidance_fn_closed_form(x2, a, shape, ori_shape, rescaler, w_energy=we, w_init=wi, init_u=u2)
        assert got.shape == want.shape and rel_l2(got, want) < 1e-6, (control, we, wi, rel_l2(got, want))
    a = _args(False, False)
    a.w_energy, a.w_init = 0.5, 0.1
    d_cf = inf.make_design_fn(a, shape, ori_shape, rescaler, closed_form=True)(x2, init_u=u2)
    d_ag = inf.make_design_fn(a, shape, ori_shape, rescaler)(x2.clone().requires_grad_(), init_u=u2)
    assert rel_l2(d_cf, d_ag) < 1e-6


def test_cascade_pipeline_matches_reference(inf):
    s = ref_loader.smoke()
    gold = torch.load(os.path.join(GOLD, "smoke_cascade_pipeline.pt"))
    torch.manual_seed(0)
    mb = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).eval()
    torch.manual_seed(0)
    ms = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=82).eval()
    shape, ori_shape = [[18, 34, 34], [18, 66, 66]], [[32, 64, 64], [32, 128, 128]]
    rescaler = torch.linspace(0.5, 3.0, 82).reshape(1, 1, 82, 1, 1)
    kw = dict(image_size=40, frames=24, timesteps=1000, sampling_timesteps=gold["steps"], ddim_sampling_eta=1.0,
              standard_fixed_ratio=100.0)
    gb = s.GaussianDiffusion(mb, rescaler[:, :, 40:], True, True, True, False, "bior1.3", "zero", shape[0], ori_shape[0], **kw)
    gs = s.GaussianDiffusion(ms, rescaler, True, True, True, True, "bior1.3", "zero", shape, ori_shape, **kw)
    args = _args(True, True)
    pipe = inf.InferencePipeline([gb, gs], args=dict(design_fn=inf.make_design_fn(args, shape, ori_shape, rescaler),
                                                     design_guidance="standard"), RESCALER=rescaler, args_general=args)
    gen = torch.Generator().manual_seed(gold["input_seed"])
    state = 0.5 * torch.randn(1, 32, 6, 128, 128, generator=gen)
    with patched_randn(NoiseTape(gold["tape_seed"])), torch.no_grad():
        outs = pipe.run_model(state)
    for o, shp, sub in zip(outs, gold["out_shape"], gold["out_sub"]):
        assert tuple(o.shape) == shp
        assert rel_l2(o.reshape(-1)[::gold["stride"]], sub) < 1e-5


def _burgers_setup(gold, GaussianDiffusion, Unet2D, dev="cpu"):
    args = types.SimpleNamespace(is_wavelet=True, pad_mode="periodization", wave_type="bior2.4", is_super_model=True,
                                 upsample_x=1, upsample_t=1, is_condition_f=True, is_condition_u0=True)
    torch.manual_seed(0)
    mb = Unet2D(dim=64, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).eval()
    torch.manual_seed(0)
    ms = Unet2D(dim=64, dim_mults=[1, 2, 4, 8], channels=17, out_dim=17, resnet_block_groups=1).eval()
    R = torch.linspace(0.5, 2.0, 17).reshape(1, 17, 1, 1).to(dev)
    kw = dict(is_wavelet=True, pad_mode="periodization", wave_type="bior2.4", timesteps=1000,
              sampling_timesteps=gold["steps"], ddim_sampling_eta=gold["eta"], is_condition_u0=True, is_condition_f=True)
    gb = GaussianDiffusion(mb.to(dev), seq_length=(64, 64), padded_shape=[41, 60], ori_shape=[81, 120],
                           loss_layer_weight=R[:, 8:17], **kw).to(dev)
    gs = GaussianDiffusion(ms.to(dev), seq_length=(128, 128), padded_shape=[[81, 120]], ori_shape=[[161, 240]],
                           is_super_model=True, upsample_t=1, upsample_x=1, loss_layer_weight=R, **kw).to(dev)
    gen = torch.Generator().manual_seed(gold["input_seed"])
    B = 2
    u_t = [torch.randn(B, 81, 120, generator=gen).to(dev), torch.randn(B, 161, 240, generator=gen).to(dev)]
    u_c = [torch.randn(B, 64, 64, generator=gen).to(dev), torch.randn(B, 128, 128, generator=gen).to(dev)]
    fs = [torch.randn(B, 4, 64, 64, generator=gen).to(dev), torch.randn(B, 4, 128, 128, generator=gen).to(dev)]
    xg = torch.randn(B, 9, 64, 64, generator=gen).clamp(-1, 1).to(dev)
    return args, mb, ms, gb, gs, R, u_t, u_c, fs, xg


def test_burgers_glue_matches_reference(monkeypatch):
    from wdno_b200.burgers import eval_glue as G
    monkeypatch.setattr(G, "DWTInverse", T.DWTInverse)
    b = ref_loader.burgers()
    gold = torch.load(os.path.join(GOLD, "burgers_cascade.pt"))
    args, mb, ms, gb, gs, R, u_t, u_c, fs, xg = _burgers_setup(gold, b.GaussianDiffusion, b.Unet2D)
    g = G.get_nablaJ_2dconv(u_target=u_t[0], args=args, shape=[41, 60], ori_shape=[81, 120], RESCALER=R[:, 8:17],
                            wu=gold["wu"], wf=gold["wf"], condition_f=True)(xg.clone())
    assert rel_l2(g, gold["grad"]) < 1e-6
    # the J schedules against the reference's own functions
    for name in ("cosine", "sigmoid", "sigmoid_flip"):
        for t in (0, 17, 999):
            assert float(G.get_scheduler(name)(t)) == float(b.model_utils.get_scheduler(name)(t))
    with patched_randn(NoiseTape(gold["tape_seed"])), torch.no_grad():
        levels = G.run_cascade(gb, gs, args, R, u_t, u_c, fs, wu=gold["wu"], wf=gold["wf"], J_scheduler="cosine")
    assert len(levels) == 2
    for (c, u, f), want in zip(levels, gold["levels"]):
        assert (tuple(c.shape), tuple(u.shape), tuple(f.shape)) == want["shapes"]
        for got, key in ((c, "coef"), (u, "u"), (f, "f")):
            assert rel_l2(got.reshape(-1)[::gold["stride"]], want[key]) < 1e-5, key
