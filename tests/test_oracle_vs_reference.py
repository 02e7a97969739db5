"""Pins the oracle restatements against the REAL reference modules imported from /root/reference
(only possible in the build container; skipped on the GPU box, where the committed goldens take over)."""
import contextlib

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


class NoiseTape:
    """deterministic noise source shared by the reference (via patched torch.randn*) and the oracle"""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)

    def __call__(self, shape, *a, **k):
        return torch.randn(tuple(shape), generator=self.g)


@contextlib.contextmanager
def patched_randn(tape):
    orig, orig_like = torch.randn, torch.randn_like

    def randn(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        if "generator" in kw:
            return orig(*shape, **kw)
        return tape(shape)

    torch.randn, torch.randn_like = randn, (lambda t, **kw: tape(t.shape))
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig, orig_like


def _perturb_norms(m):
    with torch.no_grad():
        for n, p in m.named_parameters():
            if ".norm." in n or n.endswith("gamma") or n.endswith(".g"):
                p.add_(0.2 * torch.randn_like(p))


def test_unet3d_oracle_matches_reference():
    from oracle.unet3d import Unet3DOracle
    s = ref_loader.smoke()
    torch.manual_seed(0)
    m = s.Unet3D_with_Conv3D(dim=16, dim_mults=(1, 2, 4), channels=42).eval()
    _perturb_norms(m)
    x = torch.randn(2, 24, 42, 40, 40)
    t = torch.tensor([3, 977])
    with torch.no_grad():
        ref = m(x, t)
        got = Unet3DOracle(m.state_dict())(x, t)
    assert torch.allclose(ref, got, atol=1e-5, rtol=1e-5)


def _smoke_pair(S, eta, control=True, T=1000):
    from oracle.unet3d import Unet3DOracle
    s = ref_loader.smoke()
    torch.manual_seed(0)
    m = s.Unet3D_with_Conv3D(dim=16, dim_mults=(1, 2, 4), channels=42).eval()
    _perturb_norms(m)
    gd = s.GaussianDiffusion(m, torch.ones(1), control, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64],
                             image_size=40, frames=24, timesteps=T, sampling_timesteps=S, ddim_sampling_eta=eta)
    return s, m, gd, Unet3DOracle(m.state_dict())


def test_smoke_ddim_oracle_matches_reference():
    from oracle import diffusion as D
    s, m, gd, orc = _smoke_pair(S=3, eta=1.0)
    init = torch.randn(1, 24, 40, 40)
    control = torch.randn(1, 24, 16, 40, 40)
    with patched_randn(NoiseTape(7)), torch.no_grad():
        ref = gd.sample(batch_size=1, init=init, control=control)
    sch = D.schedule("sigmoid", 1000)
    assert torch.equal(sch["alphas_cumprod"], gd.alphas_cumprod)
    with torch.no_grad():
        got = D.smoke_ddim_sample(orc, sch, (1, 24, 42, 40, 40), 3, 1.0, NoiseTape(7), [18, 34, 34], init, control)
    assert torch.allclose(ref, got, atol=2e-5, rtol=1e-5)


def test_smoke_ddpm_and_losses_oracle_match_reference():
    from oracle import diffusion as D
    s, m, gd, orc = _smoke_pair(S=None, eta=0.0, T=4)
    init = torch.randn(1, 24, 40, 40)
    control = torch.randn(1, 24, 16, 40, 40)
    with patched_randn(NoiseTape(9)), torch.no_grad():
        ref = gd.sample(batch_size=1, init=init, control=control)
    sch = D.schedule("sigmoid", 4)
    with torch.no_grad():
        got = D.smoke_ddpm_sample(orc, sch, (1, 24, 42, 40, 40), NoiseTape(9), [18, 34, 34], init, control, T=4)
    assert torch.allclose(ref, got, atol=2e-5, rtol=1e-5)
    x0 = torch.randn(2, 24, 42, 40, 40).clamp(-1, 1)
    t = torch.tensor([1, 3])
    noise = torch.randn_like(x0)
    gd.loss_layer_weight = torch.linspace(0.5, 2.0, 42).reshape(1, 1, 42, 1, 1)
    with torch.no_grad():
        ref_l = gd.p_losses(x0.clone(), t, noise.clone())
        got_l = D.smoke_p_losses(orc, sch, x0, t, noise, [18, 34, 34], gd.loss_layer_weight)
    assert abs(float(ref_l) - float(got_l)) < 1e-5 * max(1.0, abs(float(ref_l)))


def test_packing_helpers_bit_exact():
    from wdno_b200 import packing as P
    s, b = ref_loader.smoke(), ref_loader.burgers()
    # integer-tagged tensors: any permutation error changes a value
    ct = torch.arange(2 * 42 * 24 * 40 * 40, dtype=torch.float32).reshape(2, 42, 24, 40, 40)
    for ut, shape in ((None, [18, 34, 34]), ("time", [18, 34, 34]), ("space", [18, 34, 34])):
        yl_r, yh_r = s.wave_trans_2d.tensor_to_coef(ct, shape, ut)
        yl, yh = P.smoke_tensor_to_coef(ct, shape, ut)
        assert torch.equal(yl, yl_r) and list(yh) == list(yh_r) and all(torch.equal(yh[k], yh_r[k]) for k in yh)
        assert torch.equal(P.smoke_coef_to_tensor((yl, yh)), s.wave_trans_2d.coef_to_tensor((yl_r, yh_r)))
    w = torch.arange(2 * 24 * 3 * 5 * 5, dtype=torch.float32).reshape(2, 24, 3, 5, 5)
    for ty in ("time", "space"):
        assert torch.equal(P.smoke_upsample_coef(w, None, ty), s.wave_utils.upsample_coef(w, None, ty))
    bt = torch.arange(3 * 9 * 64 * 64, dtype=torch.float32).reshape(3, 9, 64, 64)
    yl_r, yh_r = b.wave_trans.tensor_to_coef(bt, [41, 60])
    yl, yh = P.burgers_tensor_to_coef(bt, [41, 60])
    assert torch.equal(yl, yl_r) and torch.equal(yh[0], yh_r[0])
    assert torch.equal(P.burgers_coef_to_tensor(yl, yh, pad=True), b.wave_trans.coef_to_tensor(yl_r, yh_r, pad=True))
    assert torch.equal(P.burgers_coef_to_tensor(yl, yh), b.wave_trans.coef_to_tensor(yl_r, yh_r))
    # multi-level replication (J = 2): coarse level repeated 2x, finest padded by its last row
    Yl2 = torch.arange(1 * 2 * 21 * 30, dtype=torch.float32).reshape(1, 2, 21, 30)
    Yh2 = [torch.arange(1 * 2 * 3 * 41 * 60, dtype=torch.float32).reshape(1, 2, 3, 41, 60),
           torch.arange(1 * 2 * 3 * 21 * 30, dtype=torch.float32).reshape(1, 2, 3, 21, 30) + 0.5]
    assert torch.equal(P.burgers_coef_to_tensor(Yl2, Yh2), b.wave_trans.coef_to_tensor(Yl2, Yh2))
    ws = torch.arange(2 * 8 * 7 * 6, dtype=torch.float32).reshape(2, 8, 7, 6)
    assert torch.equal(P.burgers_upsample_coef(ws, None), b.wavelet_utils.upsample_coef(ws, None))
    td = torch.arange(2 * 7 * 64 * 64, dtype=torch.float32).reshape(2, 7, 64, 64)
    shp = [[41, 60], [21, 30]]
    for a_, b_ in zip(P.burgers_get_wt_T(td, shp), b.wavelet_utils.get_wt_T(td, shp)):
        assert torch.equal(a_, b_)


def test_engine_state_dict_keys_match_reference():
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    s = ref_loader.smoke()
    torch.manual_seed(0)
    ref = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42)
    torch.manual_seed(0)
    mine = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    # same construction order => same default-init weights under the same seed
    assert all(torch.equal(a[k], b[k]) for k in a)


def _burgers_pair(S, eta, T=1000, dim=32, u0=True, uT=False, f=True):
    from oracle.unet2d import Unet2DOracle
    b = ref_loader.burgers()
    torch.manual_seed(0)
    m = b.Unet2D(dim=dim, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).eval()
    _perturb_norms(m)
    gd = b.GaussianDiffusion(m, seq_length=(64, 64), is_wavelet=True, pad_mode="periodization", wave_type="bior2.4",
                             padded_shape=[41, 60], ori_shape=[81, 120], timesteps=T, sampling_timesteps=S,
                             ddim_sampling_eta=eta, loss_layer_weight=torch.linspace(0.5, 2, 9).reshape(1, 9, 1, 1),
                             is_condition_u0=u0, is_condition_uT=uT, is_condition_f=f)
    return b, m, gd, Unet2DOracle(m.state_dict())


def test_unet2d_oracle_matches_reference():
    b, m, gd, orc = _burgers_pair(3, 1.0)
    x = torch.randn(2, 9, 64, 64)
    t = torch.tensor([7, 950])
    with torch.no_grad():
        assert torch.allclose(m(x, t), orc(x, t), atol=1e-5, rtol=1e-5)


def test_burgers_samplers_and_losses_oracle_match_reference():
    from oracle import diffusion as D
    b, m, gd, orc = _burgers_pair(3, 0.5, uT=True)  # eta=1 makes sqrt(1-an-sigma^2) NaN at t=999 in the reference too
    u0, uT, f = torch.randn(2, 32, 64), torch.randn(2, 32, 64), torch.randn(2, 4, 64, 64)
    with patched_randn(NoiseTape(3)), torch.no_grad():
        ref = gd.sample(batch_size=2, u_init=u0, u_final=uT, f=f)
    sch = D.schedule("cosine", 1000)
    assert torch.equal(sch["alphas_cumprod"], gd.alphas_cumprod)
    with torch.no_grad():
        got = D.burgers_ddim_sample(orc, sch, (2, 9, 64, 64), 3, 0.5, NoiseTape(3), [41, 60], u0, uT, f)
    assert torch.allclose(ref, got, atol=2e-5, rtol=1e-5)
    b, m, gd, orc = _burgers_pair(None, 0.0, T=4)
    with patched_randn(NoiseTape(4)), torch.no_grad():
        ref = gd.sample(batch_size=2, u_init=u0, f=f)
    sch = D.schedule("cosine", 4)
    with torch.no_grad():
        got = D.burgers_ddpm_sample(orc, sch, (2, 9, 64, 64), NoiseTape(4), [41, 60], u0, None, f, T=4)
    assert torch.allclose(ref, got, atol=2e-5, rtol=1e-5)
    x0 = torch.randn(2, 9, 64, 64).clamp(-1, 1)
    t = torch.tensor([0, 3])
    noise = torch.randn_like(x0)
    with torch.no_grad():
        ref_l = gd.p_losses(x0.clone(), t, noise.clone())
        got_l = D.burgers_p_losses(orc, sch, x0, t, noise, [41, 60], gd.loss_layer_weight, cond_u0=True, cond_uT=False,
                                   cond_f=True)
    assert abs(float(ref_l) - float(got_l)) < 1e-5 * max(1.0, abs(float(ref_l)))


def test_burgers_engine_state_dict_matches_reference():
    from wdno_b200.unet2d import Unet2D
    b = ref_loader.burgers()
    torch.manual_seed(0)
    ref = b.Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1)
    torch.manual_seed(0)
    mine = Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1)
    a, c = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(c.keys()) and all(torch.equal(a[k], c[k]) for k in a)
