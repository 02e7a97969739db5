"""Training step on the engine vs the REAL reference backward.
Golden: tests/golden/smoke_train_step.pt (tests/golden/make_golden.py train): loss, total gradient norm, 228 per-parameter
gradient norms and strided gradient samples of `gd.p_losses(...).backward()` of the reference smoke model (dim 64, batch 1).
Tolerances: activation gradients are fp16 (scaled), contractions fp16 x fp16 -> fp32; stated at each assert."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smoke_train_step.pt")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _model_and_diffusion():
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    torch.manual_seed(0)
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().train()
    w = torch.linspace(0.5, 2.0, 42).reshape(1, 1, 42, 1, 1)
    gd = GaussianDiffusion(m, w, True, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64], image_size=40,
                           frames=24, timesteps=1000, sampling_timesteps=250, ddim_sampling_eta=1.0).cuda()
    return m, gd


def test_wgrad_and_dgrad_kernels_vs_torch_autograd():
    """single layers: conv 3^3 (64->64, 128->64 concat), 1x1, (1,4,4) stride-2 and its transposed twin vs F.conv3d autograd"""
    import torch.nn.functional as F
    from wdno_b200.tapgemm import TapGemm
    from wdno_b200.training import ConvLayer
    g = torch.Generator().manual_seed(1)
    dev = "cuda"

    def rnd(*s):
        return torch.randn(*s, generator=g).to(dev)

    cases = [("conv", (64, 64, 3, 3, 3), (64,), (2, 6, 12, 10)), ("conv", (64, 128, 3, 3, 3), (64, 64), (1, 5, 9, 11)),
             ("conv", (128, 64, 1, 1, 1), (64,), (2, 4, 8, 8)), ("conv", (64, 48, 7, 7, 7), (48,), (1, 6, 10, 10)),
             ("down144", (64, 64, 1, 4, 4), (64,), (2, 3, 6, 8)), ("up144", (64, 64, 1, 4, 4), (64,), (2, 3, 6, 8))]
    for kind, ws, srcs, (B, D, H, W) in cases:
        weight = torch.nn.Parameter(0.05 * rnd(*ws))
        bias = torch.nn.Parameter(0.1 * rnd(ws[1] if kind == "up144" else ws[0]))
        weight.grad, bias.grad = torch.zeros_like(weight), torch.zeros_like(bias)
        cin = sum(srcs)
        fwd = TapGemm(weight, bias, kind=kind, src_channels=srcs if kind == "conv" else None, device=dev)
        layer = ConvLayer(fwd, weight, bias, kind, srcs)
        Hs, Ws = (2 * H, 2 * W) if kind == "down144" else (H, W)
        x = rnd(B, D, Hs, Ws, cin).half()
        xs = list(x.split(list(srcs), dim=-1))
        xs = [t.contiguous() for t in xs]
        xt = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
        wt, bt = weight.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
        if kind == "conv":
            yt = F.conv3d(xt, wt, bt, padding=(ws[2] // 2, ws[3] // 2, ws[4] // 2))
        elif kind == "down144":
            yt = F.conv3d(xt, wt, bt, stride=(1, 2, 2), padding=(0, 1, 1))
        else:
            yt = F.conv_transpose3d(xt, wt, bt, stride=(1, 2, 2), padding=(0, 1, 1))
        dy = rnd(*yt.shape).half()
        yt.backward(dy.float())
        dy_cl = dy.permute(0, 2, 3, 4, 1).contiguous()
        if kind == "up144":
            layer.backward_weight((dy_cl,), xs[0], 1.0)
        else:
            layer.backward_weight(tuple(xs), dy_cl, 1.0)
        assert rel_l2(weight.grad, wt.grad) < 2e-3, (kind, ws, rel_l2(weight.grad, wt.grad))
        assert rel_l2(bias.grad, bt.grad) < 2e-3, (kind, ws, "bias", rel_l2(bias.grad, bt.grad))
        off = 0
        for i, cs in enumerate(srcs):
            dx = layer.backward_input(dy_cl, i)
            want = xt.grad[:, off:off + cs].permute(0, 2, 3, 4, 1)
            assert rel_l2(dx.float(), want) < 2e-3, (kind, ws, i, rel_l2(dx.float(), want))
            off += cs


def test_groupnorm_silu_and_layernorm_backward_vs_torch_autograd():
    import torch.nn.functional as F
    from wdno_b200 import ops
    from wdno_b200.training import chan_layernorm_bwd, gn_bwd
    g = torch.Generator().manual_seed(2)
    B, D, H, W, Cc, G = 2, 3, 6, 7, 64, 8
    y = torch.randn(B, D, H, W, Cc, generator=g).cuda().half()
    gamma = (1.0 + 0.2 * torch.randn(Cc, generator=g)).cuda()
    beta = (0.1 * torch.randn(Cc, generator=g)).cuda()
    ss = (0.3 * torch.randn(B, 2 * Cc, generator=g)).cuda()
    dh = torch.randn(B, D, H, W, Cc, generator=g).cuda().half()
    # torch reference
    yt = y.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    gt, bt, st = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True), ss.clone().requires_grad_(True)
    z = F.group_norm(yt, G, gt, bt, eps=1e-5)
    z = z * (st[:, :Cc, None, None, None] + 1) + st[:, Cc:, None, None, None]
    F.silu(z).backward(dh.float().permute(0, 4, 1, 2, 3))
    # engine: statistics as the conv epilogue would produce them, affine from gn_finalize
    yf = y.float().reshape(B, D * H * W, G, Cc // G)
    stats = torch.stack((yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))), dim=-1).double().contiguous()
    count = float(D * H * W * (Cc // G))
    a, c = ops.gn_finalize(stats, gamma, beta, ss, 0, 2 * Cc, B, Cc, G, count)
    dgm, dbt, dss = torch.zeros_like(gamma), torch.zeros_like(beta), torch.zeros_like(ss)
    dy = gn_bwd(dh, y, a, c, stats, gamma, beta, dgm, dbt, G, count, 1.0, ss=ss, ss_stride=2 * Cc, d_ss=dss, dss_stride=2 * Cc)
    assert rel_l2(dy.float(), yt.grad.permute(0, 2, 3, 4, 1)) < 3e-3
    assert rel_l2(dgm, gt.grad) < 2e-3 and rel_l2(dbt, bt.grad) < 2e-3 and rel_l2(dss, st.grad) < 2e-3
    # channel LayerNorm
    for Cl in (64, 256):
        x = torch.randn(5, 9, Cl, generator=g).cuda().half()
        gm = (1.0 + 0.2 * torch.randn(Cl, generator=g)).cuda()
        dyl = torch.randn(5, 9, Cl, generator=g).cuda().half()
        xt = x.float().requires_grad_(True)
        gmt = gm.clone().requires_grad_(True)
        mean, var = xt.mean(-1, keepdim=True), xt.var(-1, unbiased=False, keepdim=True)
        ((xt - mean) / (var + 1e-5).sqrt() * gmt).backward(dyl.float())
        dg = torch.zeros_like(gm)
        dx = chan_layernorm_bwd(x, dyl, gm, dg, 1.0)
        assert rel_l2(dx.float(), xt.grad) < 3e-3 and rel_l2(dg, gmt.grad) < 2e-3


def test_p_losses_backward_reproduces_reference_golden():
    gold = torch.load(GOLD, weights_only=False)
    m, gd = _model_and_diffusion()
    sd = m.state_dict()
    assert abs(float(sum(v.double().abs().sum() for v in sd.values())) - gold["weights_checksum"]) < 1e-6 * gold["weights_checksum"]
    g = torch.Generator().manual_seed(gold["input_seed"])
    x0 = torch.randn(1, 24, 42, 40, 40, generator=g).clamp(-1, 1).cuda()
    noise = torch.randn(1, 24, 42, 40, 40, generator=g).cuda()
    t = torch.tensor([gold["t"]]).cuda()
    loss = gd.p_losses(x0, t, noise)
    assert loss.grad_fn is not None, "p_losses must be differentiable when parameters require grad"
    # forward value: fp16-operand forward vs the reference's fp32 (same bound as the inference forward)
    assert abs(float(loss) - gold["loss"]) < 5e-3 * abs(gold["loss"]), (float(loss), gold["loss"])
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
    total = float(torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())))
    report = dict(loss=float(loss), loss_ref=gold["loss"], total=total, total_ref=gold["total_norm"])
    worst_norm, worst_sub = ("", 0.0), ("", 0.0)
    for k, n in gold["grad_norms"].items():
        e = abs(float(grads[k].norm()) - n) / (n + 1e-12)
        if n > 1e-6 and e > worst_norm[1]:
            worst_norm = (k, e)
    for k, sub in gold["grad_subs"].items():
        if float(sub.norm()) > 1e-6:
            e = rel_l2(grads[k].reshape(-1)[::gold["stride"]].cpu(), sub)
            if e > worst_sub[1]:
                worst_sub = (k, e)
    report.update(worst_norm=worst_norm, worst_sub=worst_sub)
    print("TRAIN_PARITY", report)
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(report, open(os.path.join(out, "train_parity.json"), "w"), indent=1)
    # stated bounds (measured on a B200: total norm 2.2e-4, worst per-parameter norm 4.9e-4, worst strided sample 3.0e-3):
    # total gradient norm 0.2 %, every per-parameter norm 0.5 %, strided gradient samples 1.5 % rel-L2
    assert abs(total - gold["total_norm"]) < 2e-3 * gold["total_norm"], report
    assert worst_norm[1] < 5e-3, report
    assert worst_sub[1] < 1.5e-2, report


def test_two_optimizer_steps_track_the_fp32_oracle():
    """two Adam steps (clip 1.0) driven by engine gradients vs the fp32 oracle step (oracle/training.py, pinned to the reference)"""
    from oracle import diffusion as D
    from oracle import training as TR
    m, gd = _model_and_diffusion()
    g = torch.Generator().manual_seed(4)
    x0 = torch.randn(1, 24, 42, 40, 40, generator=g).clamp(-1, 1)
    noise = torch.randn(1, 24, 42, 40, 40, generator=g)
    t = torch.tensor([321])
    lr = 1e-4
    opt = torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.99))
    params = {k: v.detach().cuda().clone() for k, v in m.state_dict().items()}
    ost = TR.adam_init({k: params[k] for k in TR.trainable_names(params)})
    sch = {k: v.cuda() for k, v in D.schedule("sigmoid", 1000).items()}
    w = gd.loss_layer_weight.cuda()
    for step in range(2):
        loss = gd.p_losses(x0.cuda(), t.cuda(), noise.clone().cuda())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        lo, _, params = TR.smoke_train_step(params, ost, sch, x0.cuda(), t.cuda(), noise.clone().cuda(), [18, 34, 34], w, lr)
        assert abs(float(loss) - float(lo)) < 1e-2 * abs(float(lo)), (step, float(loss), float(lo))
    # parameters after two steps: the UPDATE (lr-sized) must agree, measured relative to the update itself
    sd0 = _model_and_diffusion()[0].state_dict()
    num = den = 0.0
    for k, p in m.named_parameters():
        if not p.requires_grad:
            continue
        du, dw = p.detach() - sd0[k], params[k] - sd0[k]
        num += float(((du - dw).double() ** 2).sum())
        den += float((dw.double() ** 2).sum())
    assert (num / den) ** 0.5 < 0.15, (num / den) ** 0.5   # Adam normalises by sqrt(v): sign-level agreement of fp16-noise gradients


def test_fused_trainer_step_matches_torch_clip_adam_on_the_same_gradients():
    """FusedTrainer (flat buffers, one clip + Adam + EMA launch) vs torch.nn.utils.clip_grad_norm_ + torch.optim.Adam fed with
    the SAME gradient values, two steps (Adam normalises by sqrt(v), so independently recomputed gradients whose near-zero
    entries differ in fp32-atomic order would not be a test of the optimiser); EMA follows ema_pytorch's schedule."""
    from wdno_b200.trainer import FusedTrainer
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(1, 24, 42, 40, 40, generator=g).clamp(-1, 1).cuda()
    ma, gda = _model_and_diffusion()
    mb, _ = _model_and_diffusion()
    tr = FusedTrainer(gda, lr=1e-4, betas=(0.9, 0.99), max_norm=1.0, ema_update_every=2, ema_update_after_step=100)
    opt = torch.optim.Adam(mb.parameters(), lr=1e-4, betas=(0.9, 0.99))
    pa, pb = dict(ma.named_parameters()), dict(mb.named_parameters())
    for step in range(2):
        tr.g.zero_()
        tr.backward(gda(x0))
        for k in pa:
            if pa[k].requires_grad:
                pb[k].grad = pa[k].grad.detach().clone()
        total = torch.nn.utils.clip_grad_norm_([p for p in mb.parameters() if p.grad is not None], 1.0)
        opt.step()
        tr.optimizer_step()
        assert abs(tr.grad_norm() - float(total)) < 1e-5 * float(total), (tr.grad_norm(), float(total))
        for k in pa:
            if pa[k].requires_grad:
                assert torch.allclose(pa[k], pb[k], atol=2e-7, rtol=1e-5), (step, k, float((pa[k] - pb[k]).abs().max()))
    ema = tr.ema_state_dict()
    for k in pa:
        if pa[k].requires_grad:
            assert torch.equal(ema[k], pa[k].detach()), k     # call 2 of update_every=2 inside the warm-up: a copy
    # the engine must see the updated weights: a fresh model loaded with the trained weights gives the same forward
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    mc = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().eval()
    mc.load_state_dict(ma.state_dict())
    t = torch.tensor([7]).cuda()
    with torch.no_grad():
        ya, yc = ma.eval()(x0, t), mc(x0, t)
    assert torch.equal(ya, yc), rel_l2(ya, yc)


def test_attention_block_backward_kernels_vs_torch_autograd():
    """Residual(PreNorm(attention)) backward through the kernels (LayerNorm / core / projection gradients) vs fp32 torch autograd
    of the same block (wdno_b200/train3d.py keeps the torch form as the cross-check path)."""
    from wdno_b200 import train3d as T3
    from wdno_b200.training import AttnGrad
    m, _ = _model_and_diffusion()
    T3.flat_grads(m)
    e = m.engine()
    rel = m.time_rel_pos_bias.relative_attention_bias.weight
    rotary = m.init_temporal_attn.fn.fn.fn.rotary_emb.freqs.detach()
    g = torch.Generator().manual_seed(8)
    cases = [("temporal", m.downs[0][3], (1, 24, 10, 12, 64)), ("temporal", m.mid_temporal_attn, (1, 24, 5, 6, 256)),
             ("linear", m.downs[1][2], (1, 6, 20, 20, 128)), ("linear", m.ups[2][2], (2, 3, 40, 40, 64)),
             ("spatial", m.mid_spatial_attn, (1, 5, 10, 10, 256))]
    for kind, mod, shape in cases:
        a = mod.fn.fn if kind == "linear" else mod.fn.fn.fn
        x = torch.randn(*shape, generator=g).cuda().half()
        dy = torch.randn(*shape, generator=g).cuda().half()
        params = [mod.fn.norm.gamma, a.to_qkv.weight, a.to_out.weight]
        if kind == "linear":
            params.append(a.to_out.bias)
            fn = T3.linattn_block_torch
        elif kind == "temporal":
            params.append(rel)
            fn = lambda xx, gm, wq, wo, re: T3.temporal_block_torch(xx, gm, wq, wo, re, rotary)
        else:
            fn = T3.mid_spatial_block_torch
        for p in params:
            p.grad.zero_()
        dx_ref = T3._torch_block_backward(fn, x, params, dy, 1.0)
        ref = [p.grad.clone() for p in params]
        for p in params:
            p.grad.zero_()
        ag = AttnGrad(kind, mod.fn.norm.gamma, a.to_qkv, a.to_out, "cuda", rel_emb=rel)
        tables = e._rel_tables(shape[1]) if kind == "temporal" else None
        dx = ag.backward(x, dy, 1.0, tables=tables)
        assert rel_l2(dx.float(), dx_ref.float()) < 4e-3, (kind, shape, "dx", rel_l2(dx.float(), dx_ref.float()))
        for p, r, name in zip(params, ref, ("gamma", "to_qkv", "to_out", "bias/rel_emb")):
            assert rel_l2(p.grad, r) < 6e-3, (kind, shape, name, rel_l2(p.grad, r))


def test_burgers_p_losses_backward_vs_fp32_oracle_autograd():
    """Burgers Unet2D(dim=128) training step (train_diffusion.py:200-212): loss and every parameter gradient of the engine vs
    autograd through the fp32 oracle network (oracle/training.py, pinned to the reference's backward on CPU)."""
    from oracle import diffusion as D
    from oracle import training as TR
    from wdno_b200.diffusion_burgers import GaussianDiffusion
    from wdno_b200.unet2d import Unet2D
    torch.manual_seed(0)
    m = Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).cuda().train()
    lw = torch.linspace(0.5, 2.0, 9).reshape(1, 9, 1, 1)
    gd = GaussianDiffusion(m, seq_length=(64, 64), is_wavelet=True, pad_mode="periodization", wave_type="bior2.4",
                           padded_shape=[41, 60], ori_shape=[81, 120], timesteps=1000, sampling_timesteps=1000,
                           is_condition_u0=True, is_condition_f=True, loss_layer_weight=lw).cuda()
    g = torch.Generator().manual_seed(6)
    x0 = torch.randn(2, 9, 64, 64, generator=g).clamp(-1, 1).cuda()
    t = torch.tensor([5, 911]).cuda()
    noise = torch.randn(2, 9, 64, 64, generator=g).cuda()
    loss = gd.p_losses(x0, t, noise.clone())
    assert loss.grad_fn is not None
    loss.backward()
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    sch = {k: v.cuda() for k, v in D.schedule("cosine", 1000).items()}
    lo, grads = TR.burgers_loss_and_grads(params, sch, x0, t, noise.clone(), [41, 60], lw.cuda())
    assert abs(float(loss) - float(lo)) < 5e-3 * abs(float(lo)), (float(loss), float(lo))
    named = dict(m.named_parameters())
    tot_e = float(torch.sqrt(sum((named[k].grad.double() ** 2).sum() for k in grads if k in named)))
    tot_o = float(torch.sqrt(sum((grads[k].double() ** 2).sum() for k in grads if k in named)))
    assert abs(tot_e - tot_o) < 5e-3 * tot_o, (tot_e, tot_o)
    worst = ("", 0.0)
    for k, go in grads.items():
        if k in named and float(go.norm()) > 1e-6 * tot_o:
            e = rel_l2(named[k].grad, go)
            if e > worst[1]:
                worst = (k, e)
    print("BURGERS_TRAIN_PARITY", dict(loss=float(loss), loss_ref=float(lo), total=tot_e, total_ref=tot_o, worst=worst))
    assert worst[1] < 3e-2, worst
