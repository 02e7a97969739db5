"""Generate the committed golden vectors by running the REAL reference modules (imported from /root/reference).

    python tests/golden/make_golden.py        # build container only

Weights are NOT stored (95 MB): the reference module built under torch.manual_seed(0) is bit-reproduced by
wdno_b200's own module under the same seed (tests/test_oracle_vs_reference.py::test_engine_state_dict_keys_match_reference),
and each fixture records a checksum of the weights/inputs it was generated with so a drifted RNG is detected, not
silently compared.  Outputs are stored as strided sub-samples (stride 7) plus norms.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader  # noqa: E402
from tests.test_oracle_vs_reference import NoiseTape, patched_randn  # noqa: E402

STRIDE = 7


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def sub(t, stride=STRIDE):
    return t.reshape(-1)[::stride].clone()


def smoke_fixture():
    s = ref_loader.smoke()
    torch.manual_seed(0)
    m = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 24, 42, 40, 40, generator=g)
    t = torch.tensor([321])
    with torch.no_grad():
        y = m(x, t)
    out = dict(weights_checksum=checksum(m.state_dict()), x_checksum=float(x.double().abs().sum()), t=t, y_sub=sub(y),
               y_norm=float(y.norm()), stride=STRIDE)
    torch.save(out, os.path.join(HERE, "smoke_unet3d_fwd.pt"))
    print("smoke_unet3d_fwd", out["weights_checksum"], out["y_norm"])
    # DDIM, 4 steps, eta = 1, base simulation config (C3 shape), injected noise
    gd = s.GaussianDiffusion(m, torch.ones(1), True, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64],
                             image_size=40, frames=24, timesteps=1000, sampling_timesteps=4, ddim_sampling_eta=1.0)
    init = torch.randn(1, 24, 40, 40, generator=g)
    control = torch.randn(1, 24, 16, 40, 40, generator=g)
    with patched_randn(NoiseTape(11)), torch.no_grad():
        smp = gd.sample(batch_size=1, init=init, control=control)
    out = dict(weights_checksum=checksum(m.state_dict()), init_checksum=float(init.double().abs().sum()),
               sample_sub=sub(smp), sample_norm=float(smp.norm()), stride=STRIDE, steps=4, eta=1.0, tape_seed=11)
    torch.save(out, os.path.join(HERE, "smoke_ddim4.pt"))
    print("smoke_ddim4", out["sample_norm"])


def burgers_fixture():
    b = ref_loader.burgers()
    torch.manual_seed(0)
    m = b.Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).eval()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 9, 64, 64, generator=g)
    t = torch.tensor([17, 803])
    with torch.no_grad():
        y = m(x, t)
    out = dict(weights_checksum=checksum(m.state_dict()), x_checksum=float(x.double().abs().sum()), t=t, y=y.clone(),
               y_norm=float(y.norm()))
    # config C1: DDIM (4 steps here), eta = 0.5 (eta = 1 gives sqrt of a rounding-negative number at t=999 in the
    # reference itself), u0 + f conditioning, padded 41x60 coefficients
    gd = b.GaussianDiffusion(m, seq_length=(64, 64), is_wavelet=True, pad_mode="periodization", wave_type="bior2.4",
                             padded_shape=[41, 60], ori_shape=[81, 120], timesteps=1000, sampling_timesteps=4,
                             ddim_sampling_eta=0.5, is_condition_u0=True, is_condition_f=True)
    u0 = torch.randn(2, 32, 64, generator=g)
    f = torch.randn(2, 4, 64, 64, generator=g)
    with patched_randn(NoiseTape(12)), torch.no_grad():
        smp = gd.sample(batch_size=2, u_init=u0, f=f)
    out.update(u0_checksum=float(u0.double().abs().sum()), sample=smp.clone(), sample_norm=float(smp.norm()), steps=4,
               eta=0.5, tape_seed=12)
    torch.save(out, os.path.join(HERE, "burgers_unet2d_ddim4.pt"))
    print("burgers", out["weights_checksum"], out["y_norm"], out["sample_norm"])


def pipeline_args(control, super_model):
    """argparse namespace of inference_2d.py reduced to what guidance_fn / InferencePipeline read"""
    import types
    return types.SimpleNamespace(is_wavelet=True, wave_type="bior1.3", pad_mode="zero", is_condition_control=control,
                                 is_condition_pad=True, is_super_model=super_model, upsample=1 if super_model else 0,
                                 image_size=64, device="cpu", w_energy=0.5, w_init=0.1)


def pipeline_inputs(gen, B, nt, n):
    """physical fields [B,nt,6,n,n] (rho, v1, v2, c1, c2, smoke-out), O(1).  The pipeline strides them down to the base
    resolution: control mode [B,256,6,64,64] -> ::8 in time; simulation mode [B,32,6,128,128] -> ::2 in space"""
    return 0.5 * torch.randn(B, nt, 6, n, n, generator=gen)


def smoke_guided_fixture():
    """rows f-1 / C5: the REAL reference guidance_fn + InferencePipeline.run_model (base model, control NOT conditioned,
    design_guidance='standard', standard_fixed_ratio=100, w_init=0.1: scripts/smoke/inf_base_control.sh) with the
    wavelet packages replaced by oracle/wavelets_torch.py"""
    s = ref_loader.smoke()
    inf = ref_loader.smoke_inference()
    torch.manual_seed(0)
    m = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).eval()
    shape, ori_shape = [18, 34, 34], [32, 64, 64]
    rescaler = torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1)
    S = 3
    gd = s.GaussianDiffusion(m, rescaler, False, True, True, False, "bior1.3", "zero", shape, ori_shape, image_size=40,
                             frames=24, timesteps=1000, sampling_timesteps=S, ddim_sampling_eta=1.0,
                             standard_fixed_ratio=100.0)
    args = pipeline_args(False, False)
    gen = torch.Generator().manual_seed(21)
    state = pipeline_inputs(gen, 1, 256, 64)
    # single gradient evaluation on a fixed point
    xg = torch.randn(1, 24, 42, 40, 40, generator=gen).clamp(-1, 1).requires_grad_()
    init_u = state[:, 0, 0]
    g = inf.guidance_fn(xg, args, shape, ori_shape, rescaler, w_energy=0.5, w_init=0.1, init_u=init_u)
    # the closure of load_model (needs checkpoints there), restated
    design_fn = lambda x, low=None, init=None, init_u=None: inf.guidance_fn(
        x, args, shape, ori_shape, rescaler, w_energy=args.w_energy, w_init=args.w_init, low=low, init=init, init_u=init_u)
    pipe = inf.InferencePipeline([gd], args=dict(design_fn=design_fn, design_guidance="standard"), RESCALER=rescaler,
                                 results_path="/tmp/wdno_golden_results", args_general=args)
    with patched_randn(NoiseTape(31)), torch.no_grad():
        out = pipe.run_model(state)
    res = dict(weights_checksum=checksum(m.state_dict()), state_checksum=float(state.double().abs().sum()),
               grad_sub=sub(g, 23), grad_norm=float(g.norm()), out_sub=sub(out, 23), out_norm=float(out.norm()),
               out_shape=tuple(out.shape), stride=23, steps=S, tape_seed=31, input_seed=21)
    torch.save(res, os.path.join(HERE, "smoke_guided_pipeline.pt"))
    print("smoke_guided", res["grad_norm"], res["out_norm"], res["out_shape"])


def smoke_cascade_fixture():
    """rows f-2 / C4: the REAL InferencePipeline.run_model with [base, super] models (is_condition_control=True,
    upsample=1, 'space'): base DDIM -> nearest x2 coefficients -> 82-channel model on [1,24,82,80,80] -> inverse
    transforms at both resolutions.  Guidance active in both (w_init=0.1, ratio 100)."""
    s = ref_loader.smoke()
    inf = ref_loader.smoke_inference()
    torch.manual_seed(0)
    mb = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).eval()
    torch.manual_seed(0)
    ms = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=82).eval()
    shape, ori_shape = [[18, 34, 34], [18, 66, 66]], [[32, 64, 64], [32, 128, 128]]
    rescaler = torch.linspace(0.5, 3.0, 82).reshape(1, 1, 82, 1, 1)
    S = 2
    kw = dict(image_size=40, frames=24, timesteps=1000, sampling_timesteps=S, ddim_sampling_eta=1.0,
              standard_fixed_ratio=100.0)
    gb = s.GaussianDiffusion(mb, rescaler[:, :, 40:], True, True, True, False, "bior1.3", "zero", shape[0], ori_shape[0], **kw)
    gs = s.GaussianDiffusion(ms, rescaler, True, True, True, True, "bior1.3", "zero", shape, ori_shape, **kw)
    args = pipeline_args(True, True)

    def design_fn(x, low=None, init=None, init_u=None):  # load_model's closure (inference_2d.py:82-91)
        import math
        if low is not None:
            up = int(math.log2(low.shape[-1] / 40))
            return inf.guidance_fn(x, args, shape[up], ori_shape[up], rescaler, w_energy=args.w_energy,
                                   w_init=args.w_init, low=low, init=init, init_u=init_u)
        return inf.guidance_fn(x, args, shape[0], ori_shape[0], rescaler[:, :, 40:], w_energy=args.w_energy,
                               w_init=args.w_init, low=low, init=init, init_u=init_u)
    pipe = inf.InferencePipeline([gb, gs], args=dict(design_fn=design_fn, design_guidance="standard"),
                                 RESCALER=rescaler, results_path="/tmp/wdno_golden_results", args_general=args)
    gen = torch.Generator().manual_seed(22)
    state = pipeline_inputs(gen, 1, 32, 128)
    with patched_randn(NoiseTape(32)), torch.no_grad():
        outs = pipe.run_model(state)
    res = dict(weights_checksum=checksum(mb.state_dict()), super_weights_checksum=checksum(ms.state_dict()),
               state_checksum=float(state.double().abs().sum()), stride=23, steps=S, tape_seed=32, input_seed=22,
               out_sub=[sub(o, 23) for o in outs], out_norm=[float(o.norm()) for o in outs],
               out_shape=[tuple(o.shape) for o in outs])
    torch.save(res, os.path.join(HERE, "smoke_cascade_pipeline.pt"))
    print("smoke_cascade", res["out_norm"], res["out_shape"])


def burgers_ref_guidance_loss():
    """burgers/ddpm_burgers/test_util.py::ddpm_guidance_loss, exec'd from the reference source text alone (the module
    itself imports the dataset / trainer / solver stack)"""
    import re
    src = open(os.path.join(ref_loader.REF_ROOT, "burgers", "ddpm_burgers", "test_util.py")).read()
    m = re.search(r"^def ddpm_guidance_loss\(.*?(?=^# Loading dataset)", src, re.S | re.M)
    ns = {"torch": torch}
    exec(m.group(0), ns)
    return ns["ddpm_guidance_loss"]


def burgers_cascade_fixture():
    """rows f-1/f-2, Burgers: guided base DDIM (nablaJ through the inverse 2-D bior2.4 'periodization' transform,
    J_scheduler='cosine') -> coefficient up-sampling -> 17-channel super model on 128x128 with `low` -> inverse
    transforms, assembled from the REAL reference pieces (GaussianDiffusion, Unet2D, wave_trans.tensor_to_coef[_super],
    wavelet_utils.upsample_coef, model_utils.get_nablaJ/get_scheduler, test_util.ddpm_guidance_loss) in the order of
    eval_ddpm_burgers.py:108-143,151-193,279-338; targets / conditions are synthetic tensors instead of dataset rows."""
    import types
    b = ref_loader.burgers()
    ref_loader.install_wavelet_shims()
    from oracle import wavelets_torch as wt
    gloss = burgers_ref_guidance_loss()
    mu = b.model_utils
    args = types.SimpleNamespace(is_wavelet=True, pad_mode="periodization", wave_type="bior2.4", is_super_model=True,
                                 upsample_x=1, upsample_t=1, is_condition_f=True, is_condition_u0=True)
    torch.manual_seed(0)
    mb = b.Unet2D(dim=64, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).eval()
    torch.manual_seed(0)
    ms = b.Unet2D(dim=64, dim_mults=[1, 2, 4, 8], channels=17, out_dim=17, resnet_block_groups=1).eval()
    S = 3
    R = torch.linspace(0.5, 2.0, 17).reshape(1, 17, 1, 1)
    Rb = R[:, 8:17]
    kw = dict(is_wavelet=True, pad_mode="periodization", wave_type="bior2.4", timesteps=1000, sampling_timesteps=S,
              ddim_sampling_eta=0.5, is_condition_u0=True, is_condition_f=True)
    gb = b.GaussianDiffusion(mb, seq_length=(64, 64), padded_shape=[41, 60], ori_shape=[81, 120], loss_layer_weight=Rb, **kw)
    gs = b.GaussianDiffusion(ms, seq_length=(128, 128), padded_shape=[[81, 120]], ori_shape=[[161, 240]],
                             is_super_model=True, upsample_t=1, upsample_x=1, loss_layer_weight=R, **kw)
    gen = torch.Generator().manual_seed(23)
    B = 2
    u_t = [torch.randn(B, 81, 120, generator=gen), torch.randn(B, 161, 240, generator=gen)]
    u_c = [torch.randn(B, 64, 64, generator=gen), torch.randn(B, 128, 128, generator=gen)]
    fs = [torch.randn(B, 4, 64, 64, generator=gen), torch.randn(B, 4, 128, 128, generator=gen)]
    wu, wf = 5.0, 0.01

    def loss_fn_of(u_target, shape, ori_shape, Rk, is_super):
        def loss_fn(x):
            x = x[:, :8] * Rk[:, :8] if is_super else x * Rk
            Yl, Yh = b.wave_trans.tensor_to_coef(x, shape)
            u_f = wt.DWTInverse(mode="periodization", wave="bior2.4")((Yl, Yh))[:, :, :ori_shape[-2], :ori_shape[-1]]
            return gloss(u_target[:, :ori_shape[-2], :ori_shape[-1]], u_f[:, 0], u_f[:, 1, :ori_shape[-2] - 1],
                         wu=wu if not is_super else 0, wf=wf if not is_super else 0, condition_f=True)
        return loss_fn

    def fields(x, shape, ori_shape, sup):
        Yl, Yh = (b.wave_trans.tensor_to_coef_super if sup else b.wave_trans.tensor_to_coef)(x, shape)
        u_f = wt.DWTInverse(mode="periodization", wave="bior2.4")((Yl, Yh))[:, :, :ori_shape[-2], :ori_shape[-1]]
        return x[:, :, :shape[-2], :shape[-1]][:, :8], u_f[:, 0], u_f[:, 1, :ori_shape[-2] - 1]
    # fixed-point gradient
    xg = torch.randn(B, 9, 64, 64, generator=gen).clamp(-1, 1)
    g = mu.get_nablaJ(loss_fn_of(u_t[0], [41, 60], [81, 120], Rb, False))(xg.clone())
    sched = mu.get_scheduler("cosine")
    with patched_randn(NoiseTape(33)), torch.no_grad():
        x = gb.sample(batch_size=B, J_scheduler=sched, u_init=u_c[0][:, :32] / Rb.squeeze()[-1],
                      u_final=u_c[0][:, -32:] / Rb.squeeze()[-1], f=fs[0] / Rb[:, 4:8], x_gt=None,
                      nablaJ=mu.get_nablaJ(loss_fn_of(u_t[0], [41, 60], [81, 120], Rb, False))) * Rb
        c0, u0, f0 = fields(x, [41, 60], [81, 120], False)
        low = b.wavelet_utils.upsample_coef(c0, [81, 120])
        low = torch.nn.functional.pad(low, (0, 128 - low.shape[-1], 0, 128 - low.shape[-2]), "constant", 0) / R[:, 8:16]
        x1 = gs.sample(batch_size=B, N_upsample=1, J_scheduler=sched, low=low, u_init=u_c[1][:, :64] / R.squeeze()[-1],
                       u_final=u_c[1][:, -64:] / R.squeeze()[-1], f=fs[1] / R[:, 4:8], x_gt=None,
                       nablaJ=mu.get_nablaJ(loss_fn_of(u_t[1], [81, 120], [161, 240], R, True))) * R
        c1, u1, f1 = fields(x1, [81, 120], [161, 240], True)
    res = dict(base_checksum=checksum(mb.state_dict()), super_checksum=checksum(ms.state_dict()), input_seed=23,
               tape_seed=33, steps=S, eta=0.5, wu=wu, wf=wf, grad=g.clone(), stride=5,
               levels=[dict(coef=sub(c, 5), u=sub(u, 5), f=sub(f, 5), shapes=(tuple(c.shape), tuple(u.shape), tuple(f.shape)),
                            u_norm=float(u.norm())) for c, u, f in ((c0, u0, f0), (c1, u1, f1))])
    torch.save(res, os.path.join(HERE, "burgers_cascade.pt"))
    print("burgers_cascade", float(g.norm()), [l["u_norm"] for l in res["levels"]], [l["shapes"] for l in res["levels"]])


# ------------------------------------------------------------------ offline coefficient builders
BUILDER_STRIDE = 3


def builder_inputs_smoke(n_sims=2, T=16, H=16, seed=77):
    """synthetic raw simulations in the reference's .npy layouts (wave_trans_2d.py:99-108): Density [H,W,1,T],
    Velocity / Control [H,W,2,T], Smoke [T,2] (positive)"""
    g = torch.Generator().manual_seed(seed)
    sims = []
    for _ in range(n_sims):
        sims.append(dict(Density=torch.randn(H, H, 1, T, generator=g).numpy(), Velocity=torch.randn(H, H, 2, T, generator=g).numpy(),
                         Control=torch.randn(H, H, 2, T, generator=g).numpy(),
                         Smoke=(torch.rand(T, 2, generator=g) + 0.1).numpy()))
    return sims


def builder_inputs_burgers(N=3, seed=78):
    g = torch.Generator().manual_seed(seed)
    return dict(u=torch.randn(N, 81, 120, generator=g), f=torch.randn(N, 80, 120, generator=g))


def write_smoke_sims(root, sims):
    import numpy as np
    for i, sim in enumerate(sims):
        d = os.path.join(root, "data", "2d", "train", "sim_{:06d}".format(i))
        os.makedirs(d, exist_ok=True)
        for k, v in sim.items():
            np.save(os.path.join(d, k + ".npy"), v)


def summarise(rec):
    """strided sub-samples + norms of every tensor of a saved record (files are MBs; the layout is what is pinned)"""
    out = {}
    for k, v in rec.items():
        if isinstance(v, list) and v and torch.is_tensor(v[0]):
            out[k] = [dict(shape=tuple(t.shape), sub=sub(t, BUILDER_STRIDE), norm=float(t.double().norm())) for t in v]
        elif isinstance(v, list):
            out[k] = [tuple(x) for x in v]
            out[k + "_type"] = type(v[0]).__name__
        else:
            out[k] = tuple(v)
            out[k + "_type"] = type(v).__name__
    return out


def builders_fixture():
    """run the reference's own __main__ builders (unchanged, wavelet stand-ins installed) on synthetic raw data in a
    temporary working directory and keep summaries of the files they write"""
    import tempfile
    res = dict(stride=BUILDER_STRIDE, smoke_seed=77, burgers_seed=78)
    with tempfile.TemporaryDirectory() as tmp:
        sims = builder_inputs_smoke()
        write_smoke_sims(tmp, sims)
        err = ref_loader.run_reference_main("smoke/wave_trans_2d.py", tmp)
        assert isinstance(err, FileNotFoundError) and "sim_000002" in str(err), err  # 20 000 ids are hard-coded
        wave_dir = os.path.join(tmp, "data", "2d", "train", "bior1.3_zero")
        res["smoke"] = {kind: [summarise(torch.load(os.path.join(wave_dir, kind + "_downsample", "{:06d}".format(i)),
                                                    weights_only=False)) for i in range(len(sims))]
                        for kind in ("time", "space")}
        os.makedirs(os.path.join(tmp, "data", "1d"), exist_ok=True)
        torch.save(builder_inputs_burgers(), os.path.join(tmp, "data", "1d", "train"))
        err = ref_loader.run_reference_main("burgers/wave_trans.py", tmp)
        assert err is None, err
        res["burgers"] = summarise(torch.load(os.path.join(tmp, "data", "1d", "coef_bior2.4_periodization_super"),
                                              weights_only=False))
    torch.save(res, os.path.join(HERE, "coef_builders.pt"))
    print("coef_builders", [c["shape"] for c in res["smoke"]["time"][0]["coef"]],
          [c["shape"] for c in res["smoke"]["space"][0]["coef"]], res["smoke"]["time"][0]["shape"],
          [c["shape"] for c in res["burgers"]["coef"]], res["burgers"]["shape"], res["burgers"]["ori_shape"])


def train_step_fixture():
    """loss + gradients of ONE training step of the real reference (smoke base model, dim 64, batch 1): `loss = gd.p_losses(...)`,
    `loss.backward()`, `clip_grad_norm_(1.0)` (diffusion_2d.py:1278-1287).  The parity gate of SURVEY section 8 row f-3."""
    s = ref_loader.smoke()
    torch.manual_seed(0)
    m = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).train()
    w = torch.linspace(0.5, 2.0, 42).reshape(1, 1, 42, 1, 1)
    gd = s.GaussianDiffusion(m, w, True, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64],
                             image_size=40, frames=24, timesteps=1000, sampling_timesteps=250, ddim_sampling_eta=1.0)
    seed, t = 21, 433
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(1, 24, 42, 40, 40, generator=g).clamp(-1, 1)
    noise = torch.randn(1, 24, 42, 40, 40, generator=g)
    wsum = checksum(m.state_dict())
    loss = gd.p_losses(x0.clone(), torch.tensor([t]), noise.clone())
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    total = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    keys = ["init_conv.weight", "init_temporal_attn.fn.fn.to_qkv.weight", "time_mlp.1.weight", "downs.0.0.block1.proj.weight",
            "downs.0.0.mlp.1.weight", "downs.1.2.fn.fn.to_qkv.weight", "mid_block1.block2.norm.weight", "mid_spatial_attn.fn.fn.to_out.weight",
            "ups.1.0.res_conv.weight", "ups.1.4.weight", "final_conv.1.weight", "time_rel_pos_bias.relative_attention_bias.weight"]
    keys = [k for k in keys if k in grads]
    out = dict(weights_checksum=wsum, input_seed=seed, t=t, loss=float(loss), total_norm=float(total), stride=STRIDE,
               grad_norms={k: float(v.norm()) for k, v in grads.items()}, grad_subs={k: sub(grads[k]) for k in keys})
    torch.save(out, os.path.join(HERE, "smoke_train_step.pt"))
    print("smoke_train_step", out["loss"], out["total_norm"], len(grads), keys)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_step_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "builders":
        builders_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "burgers_cascade":
        burgers_cascade_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "guided":
        smoke_guided_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cascade":
        smoke_cascade_fixture()
        sys.exit(0)
    if len(sys.argv) < 2 or sys.argv[1] == "smoke":
        smoke_fixture()
    if len(sys.argv) < 2 or sys.argv[1] == "burgers":
        burgers_fixture()
