"""GPU parity of the sampling glue (SURVEY.md section 8 rows f-1 guided sampling, f-2 super-resolution cascade) against
goldens produced by the REAL reference `smoke/inference_2d.py` (guidance_fn + InferencePipeline.run_model, real
reference U-Nets and GaussianDiffusion, wavelet packages stood in by oracle/wavelets_torch.py) --
tests/golden/make_golden.py `guided` / `cascade`."""
import math
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def _tape(seed):
    g = torch.Generator().manual_seed(seed)
    return lambda shape, device=None: torch.randn(tuple(shape), generator=g)


def _args(control, super_model):
    return types.SimpleNamespace(is_wavelet=True, wave_type="bior1.3", pad_mode="zero", is_condition_control=control,
                                 is_condition_pad=True, is_super_model=super_model, upsample=1 if super_model else 0,
                                 image_size=64, device="cuda", w_energy=0.5, w_init=0.1)


def _model(ch):
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    torch.manual_seed(0)
    return Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=ch)


class _OracleNet(torch.nn.Module):
    """the fp32 torch oracle U-Net (oracle/unet3d.py, pinned bit-exactly to the reference module) standing in for the
    engine inside OUR sampler: isolates the sampler / condition / wavelet kernels and the glue from the fp16-operand
    error of the network, which a few-step chain from t = 999 amplifies"""

    def __init__(self, module):
        super().__init__()
        from oracle.unet3d import Unet3DOracle
        self.channels, self.self_condition = module.channels, False
        self.orc = Unet3DOracle({k: v.detach().cuda() for k, v in module.state_dict().items()})
        self._eng = type("E", (), {"launches": 0})()

    def engine(self):
        return self._eng

    def forward(self, x, t):
        torch.set_default_device("cuda")
        try:
            return self.orc(x, t).float().contiguous()
        finally:
            torch.set_default_device("cpu")


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_guided_base_pipeline_vs_reference_golden():
    """C5 shape at batch 1: control NOT conditioned, design_guidance='standard', ratio 100, w_init 0.1, w_energy 0.5"""
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.smoke import inference_2d as inf
    gold = torch.load(os.path.join(GOLD, "smoke_guided_pipeline.pt"))
    m = _model(42)
    assert abs(_checksum(m.state_dict()) - gold["weights_checksum"]) < 1e-6 * gold["weights_checksum"]
    m = m.cuda().eval()
    shape, ori_shape = [18, 34, 34], [32, 64, 64]
    rescaler = torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1).cuda()
    gd = GaussianDiffusion(m, rescaler, False, True, True, False, "bior1.3", "zero", shape, ori_shape, image_size=40,
                           frames=24, timesteps=1000, sampling_timesteps=gold["steps"], ddim_sampling_eta=1.0,
                           standard_fixed_ratio=100.0).cuda()
    args = _args(False, False)
    gen = torch.Generator().manual_seed(gold["input_seed"])
    state = 0.5 * torch.randn(1, 256, 6, 64, 64, generator=gen)
    assert abs(float(state.double().abs().sum()) - gold["state_checksum"]) < 1e-6 * gold["state_checksum"]
    xg = torch.randn(1, 24, 42, 40, 40, generator=gen).clamp(-1, 1).cuda().requires_grad_()
    # the gradient itself: fp32 transforms + adjoint kernels vs torch autograd through conv_transpose (float32 CPU)
    g = inf.guidance_fn(xg, args, shape, ori_shape, rescaler, w_energy=0.5, w_init=0.1, init_u=state[:, 0, 0].cuda())
    assert g.shape == xg.shape
    assert rel_l2(g.reshape(-1)[::gold["stride"]].cpu(), gold["grad_sub"]) < 1e-5
    assert abs(float(g.norm()) - gold["grad_norm"]) < 1e-5 * gold["grad_norm"]
    design_fn = inf.make_design_fn(args, shape, ori_shape, rescaler)
    pipe = inf.InferencePipeline([gd], args=dict(design_fn=design_fn, design_guidance="standard"), RESCALER=rescaler,
                                 args_general=args)
    gd._noise_source = _tape(gold["tape_seed"])
    out = pipe.run_model(state)
    assert tuple(out.shape) == gold["out_shape"]
    # 3 chained U-Net calls from t = 999 + inverse transform: same error mechanism and bound as the DDIM-4 golden
    assert rel_l2(out.reshape(-1)[::gold["stride"]].cpu(), gold["out_sub"]) < 3e-2
    assert abs(float(out.norm()) - gold["out_norm"]) < 1e-2 * gold["out_norm"]


def test_guided_sampling_whole_step_graph_matches_eager_callback():
    """opt-in capture of design_fn with the step (graph_design_fn): same trajectory as the default path, also when a second
    sample() call binds different init_u / init tensors (the graph is re-captured per design closure)"""
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.smoke import inference_2d as inf
    m = _model(42).cuda().eval()
    shape, ori_shape = [18, 34, 34], [32, 64, 64]
    rescaler = torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1).cuda()
    gd = GaussianDiffusion(m, rescaler, False, True, True, False, "bior1.3", "zero", shape, ori_shape, image_size=40,
                           frames=24, timesteps=1000, sampling_timesteps=4, ddim_sampling_eta=1.0,
                           standard_fixed_ratio=100.0).cuda()
    args = _args(False, False)
    args.w_energy, args.w_init = 0.5, 0.1
    design_fn = inf.make_design_fn(args, shape, ori_shape, rescaler)
    gen = torch.Generator().manual_seed(5)
    outs = {}
    for call in range(2):
        init = torch.randn(2, 24, 40, 40, generator=gen).cuda()
        init_u = torch.randn(2, 64, 64, generator=gen).cuda()
        for mode in (False, True):
            gd.graph_design_fn = mode
            gd._noise_source = _tape(31 + call)
            outs[(call, mode)] = gd.sample(batch_size=2, design_fn=design_fn, design_guidance="standard", init=init,
                                           init_u=init_u)
        a, b = outs[(call, False)], outs[(call, True)]
        assert torch.isfinite(b).all() and rel_l2(b, a) < 1e-4, (call, rel_l2(b, a))
    assert rel_l2(outs[(1, True)], outs[(0, True)]) > 1e-2   # the second call really used its own conditions
    run = next(iter(gd._runners.values()))
    assert not getattr(run, "_gg_failed", False) and len(run._gg[1]) == 2   # with / without noise, both captured


def test_closed_form_guidance_matches_autograd_on_the_kernels():
    """guidance_fn_closed_form (waverec3 -> closed-form field gradient -> waverec3_adjoint, no autograd) against the autograd
    guidance_fn on the same kernels; its math is pinned to the reference's gradient in tests/test_pipeline_cpu.py"""
    from wdno_b200.smoke import inference_2d as inf
    shape, ori_shape = [18, 34, 34], [32, 64, 64]
    rescaler = torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1).cuda()
    gen = torch.Generator().manual_seed(12)
    x = torch.randn(2, 24, 42, 40, 40, generator=gen).clamp(-1, 1).cuda()
    u = torch.randn(2, 64, 64, generator=gen).cuda()
    for control, we, wi in ((False, 0.5, 0.1), (True, 0.7, 0.2)):
        a = _args(control, False)
        want = inf.guidance_fn(x.clone().requires_grad_(), a, shape, ori_shape, rescaler, w_energy=we, w_init=wi, init_u=u)
        got = inf.guidance_fn_closed_form(x, a, shape, ori_shape, rescaler, w_energy=we, w_init=wi, init_u=u)
        assert rel_l2(got, want) < 1e-5, (control, rel_l2(got, want))


def test_super_resolution_cascade_vs_reference_golden():
    """C4 shape at batch 1: base DDIM -> x2 coefficient up-sampling -> 82-channel model on [1,24,82,80,80]
    (`low` conditioning, replicate-padded control coefficients, N_upsample=1) -> fields at 64^2 and 128^2"""
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.smoke import inference_2d as inf
    gold = torch.load(os.path.join(GOLD, "smoke_cascade_pipeline.pt"))
    mb, ms = _model(42), _model(82)
    assert abs(_checksum(mb.state_dict()) - gold["weights_checksum"]) < 1e-6 * gold["weights_checksum"]
    assert abs(_checksum(ms.state_dict()) - gold["super_weights_checksum"]) < 1e-6 * gold["super_weights_checksum"]
    mb, ms = mb.cuda().eval(), ms.cuda().eval()
    shape, ori_shape = [[18, 34, 34], [18, 66, 66]], [[32, 64, 64], [32, 128, 128]]
    rescaler = torch.linspace(0.5, 3.0, 82).reshape(1, 1, 82, 1, 1).cuda()
    kw = dict(image_size=40, frames=24, timesteps=1000, sampling_timesteps=gold["steps"], ddim_sampling_eta=1.0,
              standard_fixed_ratio=100.0)
    gb = GaussianDiffusion(mb, rescaler[:, :, 40:], True, True, True, False, "bior1.3", "zero", shape[0], ori_shape[0], **kw).cuda()
    gs = GaussianDiffusion(ms, rescaler, True, True, True, True, "bior1.3", "zero", shape, ori_shape, **kw).cuda()
    args = _args(True, True)
    design_fn = inf.make_design_fn(args, shape, ori_shape, rescaler)
    pipe = inf.InferencePipeline([gb, gs], args=dict(design_fn=design_fn, design_guidance="standard"), RESCALER=rescaler,
                                 args_general=args)
    gen = torch.Generator().manual_seed(gold["input_seed"])
    state = 0.5 * torch.randn(1, 32, 6, 128, 128, generator=gen)
    assert abs(float(state.double().abs().sum()) - gold["state_checksum"]) < 1e-6 * gold["state_checksum"]

    def run(tol, tag):
        gb._noise_source = gs._noise_source = _tape(gold["tape_seed"])
        outs = pipe.run_model(state)
        assert len(outs) == 2
        errs = []
        for o, shp, s, n in zip(outs, gold["out_shape"], gold["out_sub"], gold["out_norm"]):
            assert tuple(o.shape) == shp
            errs.append(rel_l2(o.reshape(-1)[::gold["stride"]].cpu(), s))
            assert abs(float(o.norm()) - n) < max(tol, 1e-2) * n
        print(f"cascade [{tag}] rel-L2 per level: {errs}")
        assert max(errs) < tol, (tag, errs)
    # (1) engine networks.  Two chained 2-step samplers, each starting at t = 999 where x0 = sr*x - srm1*eps multiplies
    # the fp16-level eps error by srm1 ~ 1e2 before the clamp, and the second conditioned on the first's output:
    # measured 6.8e-2 (level 1), bound 1.5e-1; the ablation below shows everything but the network is exact
    run(1.5e-1, "engine")
    # (2) same pipeline, same kernels for sampler / conditions / transforms, but fp32 oracle networks (cuDNN fp32 on the
    # GPU vs the golden's CPU fp32): measured 7.5e-5 (level 0) / 1.4e-3 (level 1) -- fp32 ROUND-OFF alone is amplified
    # ~1e3x by this chain, which is why (1) is bounded loosely; bound 5e-3
    gb.model, gs.model = _OracleNet(mb), _OracleNet(ms)
    gb.use_cuda_graph = gs.use_cuda_graph = False
    run(5e-3, "fp32 oracle nets")


def test_super_model_sample_shapes_and_conditions():
    """82-channel model: `low` occupies channels 40:80 of every intermediate state; the last DDIM step returns x0
    without re-imposing (diffusion_2d.py:897-899), so check through a 1-step DDPM-style run on T = 2"""
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    ms = _model(82).cuda().eval()
    shape, ori_shape = [[18, 34, 34], [18, 66, 66]], [[32, 64, 64], [32, 128, 128]]
    gs = GaussianDiffusion(ms, torch.ones(1), True, True, True, True, "bior1.3", "zero", shape, ori_shape, image_size=40,
                           frames=24, timesteps=2, sampling_timesteps=None).cuda()
    low = torch.randn(1, 24, 40, 80, 80, device="cuda")
    init = torch.randn(1, 24, 80, 80, device="cuda")
    control = torch.randn(1, 24, 16, 80, 80, device="cuda")
    out = gs.sample(batch_size=1, N_upsample=1, init=init, control=control, low=low)
    assert out.shape == (1, 24, 82, 80, 80)
    assert torch.equal(out[:, :, 40:80], low)              # p_sample_loop re-imposes after every step
    assert torch.equal(out[:, :18, 24:40, :68, :68], control[:, :18, :, :68, :68])   # the pad zeroing comes after
    assert torch.equal(out[:, :, 80, :68, :68], init[:, :, :68, :68])
    assert float(out[:, :, 24:40, 68:].abs().max()) == 0.0
    assert float(out[:, 18:, :24].abs().max()) == 0.0      # padded frames
    assert float(out[:, :, :24, 68:].abs().max()) == 0.0   # padded rows (coef shape 66 + 2)
    assert math.isfinite(float(out.norm()))


class _OracleNet2D(torch.nn.Module):
    """fp32 torch oracle Unet2D (oracle/unet2d.py) in place of the engine -- see _OracleNet"""

    def __init__(self, module):
        super().__init__()
        from oracle.unet2d import Unet2DOracle
        self.channels, self.self_condition = module.channels, False
        self.orc = Unet2DOracle({k: v.detach().cuda() for k, v in module.state_dict().items()})
        self._eng = type("E", (), {"launches": 0})()

    def engine(self):
        return self._eng

    def forward(self, x, t, *a):
        torch.set_default_device("cuda")
        try:
            return self.orc(x, t).float().contiguous()
        finally:
            torch.set_default_device("cpu")


def test_burgers_guided_cascade_vs_reference_golden():
    """Burgers rows f-1/f-2: nablaJ through the inverse bior2.4 'periodization' transform (adjoint kernels), cosine J
    schedule, base DDIM -> x2 coefficients -> 17-channel super model on 128x128 with `low` -> u, f at both levels"""
    from tests.test_pipeline_cpu import _burgers_setup
    from wdno_b200.burgers import eval_glue as G
    from wdno_b200.diffusion_burgers import GaussianDiffusion
    from wdno_b200.unet2d import Unet2D
    gold = torch.load(os.path.join(GOLD, "burgers_cascade.pt"))
    args, mb, ms, gb, gs, R, u_t, u_c, fs, xg = _burgers_setup(gold, GaussianDiffusion, Unet2D, dev="cuda")
    assert abs(_checksum(mb.state_dict()) - gold["base_checksum"]) < 1e-6 * gold["base_checksum"]
    assert abs(_checksum(ms.state_dict()) - gold["super_checksum"]) < 1e-6 * gold["super_checksum"]
    g = G.get_nablaJ_2dconv(u_target=u_t[0], args=args, shape=[41, 60], ori_shape=[81, 120], RESCALER=R[:, 8:17],
                            wu=gold["wu"], wf=gold["wf"], condition_f=True)(xg.clone())
    assert rel_l2(g.cpu(), gold["grad"]) < 1e-5

    def run(tol, tag):
        gb._noise_source = gs._noise_source = _tape(gold["tape_seed"])
        levels = G.run_cascade(gb, gs, args, R, u_t, u_c, fs, wu=gold["wu"], wf=gold["wf"], J_scheduler="cosine")
        errs = []
        for (c, u, f), want in zip(levels, gold["levels"]):
            assert (tuple(c.shape), tuple(u.shape), tuple(f.shape)) == want["shapes"]
            errs.append([rel_l2(t.reshape(-1)[::gold["stride"]].cpu(), want[k]) for t, k in ((c, "coef"), (u, "u"), (f, "f"))])
        print(f"burgers cascade [{tag}] rel-L2 (coef, u, f) per level: {errs}")
        assert max(max(e) for e in errs) < tol, (tag, errs)
    # two chained 3-step samplers from t = 999 (error mechanism: DESIGN.md section 3): measured 1.0e-2 .. 2.5e-2
    run(6e-2, "engine")
    # same kernels for sampler / conditions / transforms / guidance adjoint, fp32 oracle networks: measured 3e-7 .. 6e-6
    gb.model, gs.model = _OracleNet2D(mb), _OracleNet2D(ms)
    gb.use_cuda_graph = gs.use_cuda_graph = False
    run(5e-5, "fp32 oracle nets")
