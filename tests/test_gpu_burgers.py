"""GPU parity of the Burgers path (Unet2D + GaussianDiffusion of diffusion_1d.py) -- BASELINE configs C1 / C2 shapes.
Tolerances as in test_gpu_smoke.py: fp16 operands / activations with fp32 accumulation vs the fp32 reference."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _model():
    from wdno_b200.unet2d import Unet2D
    torch.manual_seed(0)
    return Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1)


def _tape(seed):
    g = torch.Generator().manual_seed(seed)
    return lambda shape, device=None: torch.randn(tuple(shape), generator=g)


def _diffusion(m, S, eta, T=1000, **kw):
    from wdno_b200.diffusion_burgers import GaussianDiffusion
    return GaussianDiffusion(m, seq_length=(64, 64), is_wavelet=True, pad_mode="periodization", wave_type="bior2.4",
                             padded_shape=[41, 60], ori_shape=[81, 120], timesteps=T, sampling_timesteps=S,
                             ddim_sampling_eta=eta, is_condition_u0=True, is_condition_f=True, **kw).cuda()


def test_unet2d_forward_and_ddim_vs_reference_golden():
    from oracle.unet2d import Unet2DOracle
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gold = torch.load(os.path.join(GOLD, "burgers_unet2d_ddim4.pt"))
    m = _model()
    ck = float(sum(v.double().abs().sum() for v in m.state_dict().values()))
    assert abs(ck - gold["weights_checksum"]) < 1e-6 * gold["weights_checksum"]
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 9, 64, 64, generator=g)
    assert abs(float(x.double().abs().sum()) - gold["x_checksum"]) < 1e-6 * gold["x_checksum"]
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(x.cuda(), gold["t"].cuda())
    assert rel_l2(y.cpu(), gold["y"]) < 5e-3        # measured ~1.5e-3
    # per-layer vs the oracle
    taps_e, taps_o = {}, {}
    with torch.no_grad():
        m.engine().forward(x.cuda(), gold["t"].cuda(), taps=taps_e)
        orc = Unet2DOracle(m.state_dict())
        orc.sd = {k: v.cuda() for k, v in orc.sd.items()}
        torch.set_default_device("cuda")
        try:
            orc(x.cuda(), gold["t"].cuda(), taps=taps_o)
        finally:
            torch.set_default_device("cpu")
    for k, vo in taps_o.items():
        assert rel_l2(taps_e[k][:, 0].permute(0, 3, 1, 2).float(), vo) < 5e-3, k
    # DDIM-4 chain (config C1 conditioning) with the fixture's noise tape, eager and CUDA-graph paths
    u0 = torch.randn(2, 32, 64, generator=g)
    f = torch.randn(2, 4, 64, 64, generator=g)
    assert abs(float(u0.double().abs().sum()) - gold["u0_checksum"]) < 1e-6 * gold["u0_checksum"]
    gd = _diffusion(m, gold["steps"], gold["eta"])
    for graph in (False, True):
        gd.use_cuda_graph = graph
        gd._noise_source = _tape(gold["tape_seed"])
        smp = gd.sample(batch_size=2, u_init=u0.cuda(), f=f.cuda()).cpu()
        assert smp.shape == (2, 9, 64, 64)
        assert rel_l2(smp, gold["sample"]) < 3e-2, graph   # t=999 start amplifies the eps error
    # conditions are exact copies / exact zeros
    assert torch.equal(smp[:, -1, :32, :60], u0[:, :, :60])
    assert torch.equal(smp[:, 4:8, :41, :60], f[:, :, :41, :60])
    assert float(smp[:, :-1, 41:].abs().max()) == 0.0 and float(smp[:, :, :, 60:].abs().max()) == 0.0


def test_burgers_ddpm_loop_guidance_and_p_losses_vs_oracle():
    from oracle import diffusion as D
    from oracle.unet2d import Unet2DOracle
    m = _model().cuda().eval()
    T = 3
    lw = torch.linspace(0.5, 2.0, 9).reshape(1, 9, 1, 1)
    gd = _diffusion(m, None, 0.0, T=T, loss_layer_weight=lw)
    orc = Unet2DOracle({k: v.cpu() for k, v in m.state_dict().items()})
    sch = D.schedule("cosine", T)
    u0, f = torch.randn(2, 32, 64), torch.randn(2, 4, 64, 64)
    gd._noise_source = _tape(6)
    got = gd.sample(batch_size=2, u_init=u0.cuda(), f=f.cuda()).cpu()
    tp = _tape(6)
    with torch.no_grad():
        want = D.burgers_ddpm_sample(orc, sch, (2, 9, 64, 64), lambda s: tp(s), [41, 60], u0, None, f, T=T)
    assert rel_l2(got, want) < 3e-2
    # guided DDIM: eps += nablaJ(x0) * J_scheduler(t) with a closed-form gradient
    gd2 = _diffusion(m, 3, 0.5)
    target = torch.randn(2, 9, 64, 64)
    nabla = lambda x0: 0.05 * (x0 - target.to(x0.device))
    sched = lambda t: 1.0 + t / 1000.0
    gd2._noise_source = _tape(8)
    got = gd2.sample(batch_size=2, u_init=u0.cuda(), f=f.cuda(), nablaJ=nabla, J_scheduler=sched).cpu()
    tp = _tape(8)
    with torch.no_grad():
        want = D.burgers_ddim_sample(orc, D.schedule("cosine", 1000), (2, 9, 64, 64), 3, 0.5, lambda s: tp(s), [41, 60], u0,
                                     None, f, guidance=lambda x0, t: 0.05 * (x0 - target) * sched(t))
    assert rel_l2(got, want) < 3e-2
    x0 = torch.randn(2, 9, 64, 64).clamp(-1, 1)
    t = torch.tensor([0, 2])
    noise = torch.randn_like(x0)
    got_l = float(gd.p_losses(x0.cuda(), t.cuda(), noise.clone().cuda()))
    with torch.no_grad():
        want_l = float(D.burgers_p_losses(orc, sch, x0, t, noise, [41, 60], lw))
    assert abs(got_l - want_l) < 1e-2 * abs(want_l)
