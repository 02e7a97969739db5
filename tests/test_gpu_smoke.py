"""GPU parity tests (B200): CUDA path through the C ABI vs the oracle / committed goldens.
Tolerances: the dense contractions use fp16 operands (11-bit significand, the same as the TF32 the reference's own
cuDNN path uses on a GPU) with fp32 accumulation and fp16 activations between layers, so per-layer / per-forward
comparisons against the fp32 oracle use relative-L2 bounds stated at each assert; pure fp32 element-wise kernels
and index permutations are compared exactly or to 1e-6."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_library_loads_on_sm100():
    from wdno_b200 import _lib
    L = _lib.lib()
    assert L.wdno_version() >= 100
    assert L.wdno_device_cc() == 100


@pytest.mark.parametrize("case", list(range(10)) + [12, 13, 14])
def test_tapgemm_against_torch_conv(case):
    import gpu_probe_tapgemm as probe
    r = probe.run_case(case)
    # fp16 output rounding (2^-11 relative to the largest magnitude) dominates
    for k, v in r.items():
        if k.startswith("err"):
            for e in (v if isinstance(v, list) else [v]):
                assert e < 1.5e-3, (r["name"], k, v)


def test_elementwise_kernels():
    from wdno_b200 import ops
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(2, 3, 42, 8, 9, device=dev)
    p = ops.pack_bfchw_f16(x, 48)
    ref = torch.zeros(2, 3, 8, 9, 48, device=dev, dtype=torch.float16)
    ref[..., :42] = x.permute(0, 1, 3, 4, 2).half()
    assert torch.equal(p, ref)
    for C in (64, 128, 256, 512, 1024):
        v = (torch.randn(3, 5, 7, C, device=dev) * 2 + 0.3).half()
        g = torch.randn(C, device=dev)
        got = ops.chan_layernorm(v, g)
        vf = v.float()
        want = (vf - vf.mean(-1, keepdim=True)) / (vf.var(-1, unbiased=False, keepdim=True) + 1e-5).sqrt() * g
        assert rel_l2(got.float(), want) < 1e-3
    B, C, G = 2, 64, 8
    y = torch.randn(B, 4, 5, 6, C, device=dev).half()
    stats = torch.zeros(B, G, 2, dtype=torch.float64, device=dev)
    yg = y.double().reshape(B, -1, G, C // G)
    stats[:, :, 0] = yg.sum(dim=(1, 3))
    stats[:, :, 1] = (yg ** 2).sum(dim=(1, 3))
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    ss = torch.randn(B, 2 * C + 10, device=dev)
    a, c = ops.gn_finalize(stats, gamma, beta, ss, 10, ss.shape[1], B, C, G, 4 * 5 * 6 * (C // G))
    r = torch.randn_like(y)
    got = ops.gn_silu_add(y, a, c, resid=r)
    yn = F.group_norm(y.float().permute(0, 4, 1, 2, 3), G, gamma, beta, eps=1e-5)
    sc, sh = ss[:, 10:10 + C], ss[:, 10 + C:10 + 2 * C]
    want = F.silu(yn * (sc[:, :, None, None, None] + 1) + sh[:, :, None, None, None]).permute(0, 2, 3, 4, 1) + r.float()
    assert rel_l2(got.float(), want) < 1e-3
    # time embedding MLP + concatenated block MLPs
    tm = torch.tensor([0.0, 17.0, 999.0], device=dev)
    w1, b1 = torch.randn(256, 64, device=dev) * 0.1, torch.randn(256, device=dev)
    w2, b2 = torch.randn(256, 256, device=dev) * 0.05, torch.randn(256, device=dev)
    emb, emb_silu = ops.time_mlp(tm, w1, b1, w2, b2)
    half = 32
    fr = torch.exp(torch.arange(half, device=dev) * -(math.log(10000) / (half - 1)))
    e = tm[:, None] * fr[None]
    sin = torch.cat((e.sin(), e.cos()), -1)
    want = F.linear(F.gelu(F.linear(sin, w1, b1)), w2, b2)
    assert torch.allclose(emb, want, atol=2e-4, rtol=2e-4)
    assert torch.allclose(emb_silu, F.silu(want), atol=2e-4, rtol=2e-4)
    wl, bl = torch.randn(300, 256, device=dev) * 0.05, torch.randn(300, device=dev)
    assert torch.allclose(ops.small_linear(emb_silu, wl, bl), F.linear(emb_silu, wl, bl), atol=2e-4, rtol=2e-4)


def test_attention_cores():
    from oracle.unet3d import rel_pos_bias, rotary
    from wdno_b200 import ops
    dev = "cuda"
    torch.manual_seed(1)
    B, Fr, H, W = 2, 24, 5, 6
    qkv = torch.randn(B, Fr, H, W, 384, device=dev).half()
    scale = 32 ** -0.5
    # temporal attention with rotary + bias
    emb = torch.randn(32, 4)
    bias = rel_pos_bias(emb, Fr).contiguous().to(dev)
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))
    ang = torch.arange(Fr, dtype=torch.float32)[:, None] * freqs[None]
    rot = (ang.cos().contiguous().to(dev), ang.sin().contiguous().to(dev))
    got = ops.softmax_attn(qkv, B * H * W, Fr, H * W, Fr * H * W, 1, H * W, scale, bias=bias, rot=rot)
    tok = qkv.float().permute(0, 2, 3, 1, 4).reshape(B, H * W, Fr, 384).cpu()
    q, k, v = [u.reshape(B, H * W, Fr, 4, 32).transpose(2, 3) for u in tok.chunk(3, -1)]
    q = rotary(q * scale, freqs)
    k = rotary(k, freqs)
    sim = q @ k.transpose(-1, -2) + bias.cpu()
    out = (sim.softmax(-1) @ v).transpose(2, 3).reshape(B, H, W, Fr, 128).permute(0, 3, 1, 2, 4)
    assert rel_l2(got.float().cpu(), out) < 2e-3
    # spatial softmax attention (no bias / rotary), n = H*W tokens per frame, n > 32 exercises key chunking
    H2, W2 = 10, 10
    qkv2 = torch.randn(B, 3, H2, W2, 384, device=dev).half()
    got2 = ops.softmax_attn(qkv2, B * 3, H2 * W2, 1, H2 * W2, 0, 1, scale)
    tok = qkv2.float().reshape(B * 3, H2 * W2, 384).cpu()
    q, k, v = [u.reshape(B * 3, H2 * W2, 4, 32).transpose(1, 2) for u in tok.chunk(3, -1)]
    out2 = (((q * scale) @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(B, 3, H2, W2, 128)
    assert rel_l2(got2.float().cpu(), out2) < 2e-3
    # the tensor-core online-softmax path (32 < n <= 512): super-resolution mid attention (n = 400), Burgers (n = 64), ragged n
    for (hh, ww) in ((20, 20), (8, 8), (7, 5), (16, 32)):
        qx = torch.randn(B, 2, hh, ww, 384, device=dev).half()
        gx = ops.softmax_attn(qx, B * 2, hh * ww, 1, hh * ww, 0, 1, scale)
        tk = qx.float().reshape(B * 2, hh * ww, 384).cpu()
        q_, k_, v_ = [u.reshape(B * 2, hh * ww, 4, 32).transpose(1, 2) for u in tk.chunk(3, -1)]
        ox = (((q_ * scale) @ k_.transpose(-1, -2)).softmax(-1) @ v_).transpose(1, 2).reshape(B, 2, hh, ww, 128)
        assert rel_l2(gx.float().cpu(), ox) < 2e-3, (hh, ww)
    # linear attention
    got3 = ops.linear_attn(qkv2, B * 3, H2 * W2, scale)
    q, k, v = [u.reshape(B * 3, H2 * W2, 4, 32).permute(0, 2, 3, 1) for u in tok.chunk(3, -1)]  # b h d n
    q = q.softmax(dim=-2) * scale
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out3 = torch.einsum("bhde,bhdn->bhen", ctx, q).permute(0, 3, 1, 2).reshape(B, 3, H2, W2, 128)
    assert rel_l2(got3.float().cpu(), out3) < 2e-3


def _seed0_model():
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    torch.manual_seed(0)
    return Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42)


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def test_unet3d_forward_vs_reference_golden_and_oracle():
    from oracle.unet3d import Unet3DOracle
    gold = torch.load(os.path.join(GOLD, "smoke_unet3d_fwd.pt"))
    m = _seed0_model()
    assert abs(_checksum(m.state_dict()) - gold["weights_checksum"]) < 1e-6 * gold["weights_checksum"], \
        "seed-0 weights differ from the ones the golden was generated with"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 24, 42, 40, 40, generator=g)
    assert abs(float(x.double().abs().sum()) - gold["x_checksum"]) < 1e-6 * gold["x_checksum"]
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(x.cuda(), gold["t"].cuda())
    got = y.reshape(-1)[::gold["stride"]].cpu()
    # fp16-operand / fp16-activation pipeline vs the reference's fp32 CPU forward: measured 1.2e-3, bound 4e-3
    assert rel_l2(got, gold["y_sub"]) < 4e-3
    assert abs(float(y.norm()) - gold["y_norm"]) < 4e-3 * gold["y_norm"]
    # per-layer against the oracle on the same device (B = 2, random t)
    x2 = torch.randn(2, 24, 42, 40, 40, device="cuda")
    t2 = torch.tensor([5, 900], device="cuda")
    taps_e, taps_o = {}, {}
    with torch.no_grad():
        y2 = m.engine().forward(x2, t2, taps=taps_e)
        orc = Unet3DOracle(m.state_dict())
        orc.sd = {k: v.cuda() for k, v in orc.sd.items()}
        torch.set_default_device("cuda")
        try:
            yo = orc(x2, t2, taps=taps_o)
        finally:
            torch.set_default_device("cpu")
    for k, vo in taps_o.items():
        if k in taps_e:
            assert rel_l2(taps_e[k].permute(0, 4, 1, 2, 3).float(), vo) < 4e-3, k
    assert rel_l2(y2, yo) < 4e-3


def test_ddim_and_ddpm_step_kernels_exact():
    from wdno_b200 import ops
    dev = "cuda"
    torch.manual_seed(2)
    B, Fr, C, H, W = 2, 6, 42, 8, 8
    x = torch.randn(B, Fr, C, H, W, device=dev)
    eps = torch.randn_like(x)
    noise = torch.randn_like(x)
    init = torch.randn(B, Fr, H, W, device=dev)
    control = torch.randn(B, Fr, 16, H, W, device=dev)
    coef = torch.tensor([1.7, 1.3, 0.8, 0.5, 0.2, 0.0, 0.0, 0.0], device=dev)
    prog = ops.CondProgram()
    prog.copy(init, "bfyx", c=(-2, -1))
    prog.copy(control, "bfcyx", c=(24, 40))
    prog.zero(f=(4, None), c=(0, -2)).zero(f=(4, None), c=(-1, None)).zero(c=(0, -1), y=(6, None)).zero(c=(0, -1), x=(7, None))
    built = prog.build(Fr, C, H, W)
    from oracle.diffusion import smoke_impose
    sr, srm1, san, cc, sg = [coef[i] for i in range(5)]
    x0 = (sr * x - srm1 * eps).clamp(-1, 1)
    e2 = (sr * x - x0) / srm1
    want = x0 * san + cc * e2 + sg * noise
    smoke_impose(want, [4, 6, 7], init, control)
    got = x.clone()
    ops.ddim_step(got, eps, noise, coef, built, 1)
    assert torch.equal(got, want)
    coef_last = coef.clone()
    coef_last[5] = 1.0
    got = x.clone()
    ops.ddim_step(got, eps, None, coef_last, built, 1)
    assert torch.equal(got, x0)  # smoke: the last step returns x0 WITHOUT re-imposing the conditions
    got = x.clone()
    ops.ddim_step(got, eps, None, coef_last, built, 2)
    assert torch.equal(got, smoke_impose(x0.clone(), [4, 6, 7], init, control))
    cp = torch.tensor([1.7, 1.3, 0.4, 0.6, 0.3, 0.0, 0.0, 0.0], device=dev)
    got = x.clone()
    ops.ddpm_step(got, eps, noise, cp, built, 2)
    want = smoke_impose(cp[2] * x0 + cp[3] * x + cp[4] * noise, [4, 6, 7], init, control)
    assert torch.equal(got, want)
    # guidance term: eps += gscale * g before x0
    gcoef = coef.clone()
    gcoef[6] = 0.25
    g = torch.randn_like(x)
    got = x.clone()
    ops.ddim_step(got, eps, noise, gcoef, built, 0, guidance=g)
    e = eps + 0.25 * g
    x0g = (sr * x - srm1 * e).clamp(-1, 1)
    want = x0g * san + cc * ((sr * x - x0g) / srm1) + sg * noise
    assert torch.equal(got, want)


def _tape(seed):
    g = torch.Generator().manual_seed(seed)
    return lambda shape, device=None: torch.randn(tuple(shape), generator=g)


def _c3_diffusion(m, S, eta=1.0, T=1000):
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    return GaussianDiffusion(m, torch.ones(1), True, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64],
                             image_size=40, frames=24, timesteps=T, sampling_timesteps=S, ddim_sampling_eta=eta).cuda()


def test_smoke_ddim_sample_vs_reference_golden():
    gold = torch.load(os.path.join(GOLD, "smoke_ddim4.pt"))
    m = _seed0_model().cuda().eval()
    gd = _c3_diffusion(m, gold["steps"], gold["eta"])
    g = torch.Generator().manual_seed(1)
    torch.randn(1, 24, 42, 40, 40, generator=g)  # the fixture drew x first
    init = torch.randn(1, 24, 40, 40, generator=g)
    control = torch.randn(1, 24, 16, 40, 40, generator=g)
    assert abs(float(init.double().abs().sum()) - gold["init_checksum"]) < 1e-6 * gold["init_checksum"]
    for graph in (False, True):
        gd.use_cuda_graph = graph
        gd._noise_source = _tape(gold["tape_seed"])
        smp = gd.sample(batch_size=1, init=init.cuda(), control=control.cuda())
        got = smp.reshape(-1)[::gold["stride"]].cpu()
        # 4 chained U-Net calls from t=999, where x0 = sr*x - srm1*eps amplifies the fp16-level (1.2e-3) eps error by
        # srm1 ~ 1e2 before the clamp: measured 1.1e-2, bound 3e-2 (numerics)
        assert rel_l2(got, gold["sample_sub"]) < 3e-2, graph
    # the conditioned channels are exact copies except where the last step skipped re-imposition
    assert smp.shape == (1, 24, 42, 40, 40)


def test_smoke_ddpm_loop_and_p_losses_vs_oracle():
    from oracle import diffusion as D
    from oracle.unet3d import Unet3DOracle
    m = _seed0_model().cuda().eval()
    T = 3
    gd = _c3_diffusion(m, None, 0.0, T=T)
    init = torch.randn(1, 24, 40, 40)
    control = torch.randn(1, 24, 16, 40, 40)
    gd._noise_source = _tape(5)
    got = gd.sample(batch_size=1, init=init.cuda(), control=control.cuda()).cpu()
    orc = Unet3DOracle(m.state_dict())
    orc.sd = {k: v.cpu() for k, v in orc.sd.items()}
    sch = D.schedule("sigmoid", T)
    tape = _tape(5)
    with torch.no_grad():
        want = D.smoke_ddpm_sample(orc, sch, (1, 24, 42, 40, 40), lambda s: tape(s), [18, 34, 34], init, control, T=T)
    assert rel_l2(got, want) < 3e-2
    x0 = torch.randn(2, 24, 42, 40, 40).clamp(-1, 1)
    t = torch.tensor([0, 2])
    noise = torch.randn_like(x0)
    lw = torch.linspace(0.5, 2.0, 42).reshape(1, 1, 42, 1, 1)
    gd.loss_layer_weight = lw
    got_l = float(gd.p_losses(x0.cuda(), t.cuda(), noise.cuda()))
    with torch.no_grad():
        want_l = float(D.smoke_p_losses(orc, sch, x0, t, noise, [18, 34, 34], lw))
    assert abs(got_l - want_l) < 1e-2 * abs(want_l)
