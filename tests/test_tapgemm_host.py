"""CPU checks of the tap-GEMM host plan (weight packing, tap tables, phases) through the numpy emulator."""
import torch
import torch.nn.functional as F

from oracle.tapgemm_emu import emulate
from wdno_b200.tapgemm import TapGemm


def cl(x):  # [B,C,D,H,W] -> fp16 channels-last [B,D,H,W,C]
    return x.permute(0, 2, 3, 4, 1).contiguous().to(torch.float16)


def uncl(y):  # [B,D,H,W,C] -> [B,C,D,H,W] fp32
    return y.permute(0, 4, 1, 2, 3).float()


def test_conv3x3x3_concat_bias_stats():
    torch.manual_seed(0)
    B, D, H, W = 1, 5, 6, 7
    x0 = torch.randn(B, 16, D, H, W).half().float()
    x1 = torch.randn(B, 16, D, H, W).half().float()
    w = (torch.randn(24, 32, 3, 3, 3) * 0.1).half().float()
    bias = torch.randn(24)
    plan = TapGemm(w, bias, src_channels=(16, 16), device="cpu", n_tile=32)
    out, stats = emulate(plan, cl(x0), cl(x1), want_stats=True, groups=3)
    ref = F.conv3d(torch.cat([x0, x1], 1), w, bias, padding=1)
    assert torch.allclose(uncl(out), ref, atol=1e-4, rtol=1e-4)
    rs = ref.reshape(B, 3, -1)
    assert torch.allclose(stats[:, :, 0].float(), rs.sum(-1), atol=1e-2)
    assert torch.allclose(stats[:, :, 1].float(), (rs ** 2).sum(-1), rtol=1e-4)


def test_conv7_init_padded_channels():
    torch.manual_seed(1)
    B, D, H, W = 1, 4, 5, 6
    x = torch.zeros(B, 48, D, H, W)
    x[:, :42] = torch.randn(B, 42, D, H, W).half().float()
    w = (torch.randn(16, 42, 7, 7, 7) * 0.05).half().float()
    plan = TapGemm(w, None, src_channels=(48,), device="cpu", kc=16)
    out, _ = emulate(plan, cl(x))
    ref = F.conv3d(x[:, :42], w, None, padding=3)
    assert torch.allclose(uncl(out), ref, atol=2e-4, rtol=1e-4)


def test_linear_1x1_resid_and_fp32_out():
    torch.manual_seed(2)
    B, D, H, W = 2, 3, 4, 5
    x = torch.randn(B, 32, D, H, W).half().float()
    w = (torch.randn(42, 32) * 0.1).half().float()
    b = torch.randn(42)
    plan = TapGemm(w, b, device="cpu")
    out, _ = emulate(plan, cl(x), out_fp32_bfchw=True)
    ref = F.conv3d(x, w[:, :, None, None, None], b)
    assert torch.allclose(out.permute(0, 2, 1, 3, 4), ref, atol=1e-4, rtol=1e-4)
    w2 = (torch.randn(32, 32) * 0.1).half().float()
    plan2 = TapGemm(w2, None, device="cpu")
    out2, _ = emulate(plan2, cl(x), resid=cl(x))
    ref2 = F.conv3d(x, w2[:, :, None, None, None]) + x
    assert torch.allclose(uncl(out2), ref2, atol=1e-4, rtol=1e-4)


def test_down144_and_up144():
    torch.manual_seed(3)
    B, D, H, W = 1, 2, 8, 8
    x = torch.randn(B, 16, D, H, W).half().float()
    wd = (torch.randn(16, 16, 1, 4, 4) * 0.1).half().float()
    bd = torch.randn(16)
    plan = TapGemm(wd, bd, kind="down144", device="cpu", n_tile=16)
    out, _ = emulate(plan, cl(x))
    ref = F.conv3d(x, wd, bd, stride=(1, 2, 2), padding=(0, 1, 1))
    assert torch.allclose(uncl(out), ref, atol=1e-4, rtol=1e-4)
    wu = (torch.randn(16, 16, 1, 4, 4) * 0.1).half().float()
    bu = torch.randn(16)
    plan_u = TapGemm(wu, bu, kind="up144", device="cpu", n_tile=16)
    out_u, _ = emulate(plan_u, cl(x))
    ref_u = F.conv_transpose3d(x, wu, bu, stride=(1, 2, 2), padding=(0, 1, 1))
    assert torch.allclose(uncl(out_u), ref_u, atol=1e-4, rtol=1e-4)


def test_burgers_unshuffle_and_up2conv_2d():
    torch.manual_seed(4)
    B, H, W = 2, 8, 8
    x = torch.randn(B, 16, H, W).half().float()
    w = (torch.randn(32, 64, 1, 1) * 0.1).half().float()
    b = torch.randn(32)
    plan = TapGemm(w, b, kind="unshuffle", device="cpu", n_tile=32)
    out, _ = emulate(plan, cl(x[:, :, None]))
    xu = x.reshape(B, 16, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, 64, 4, 4)
    ref = F.conv2d(xu, w, b)
    assert torch.allclose(uncl(out)[:, :, 0], ref, atol=1e-4, rtol=1e-4)
    w3 = (torch.randn(16, 16, 3, 3) * 0.1).half().float()
    plan3 = TapGemm(w3, None, up2=True, device="cpu", n_tile=16)
    out3, _ = emulate(plan3, cl(x[:, :, None]))
    ref3 = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w3, padding=1)
    assert torch.allclose(uncl(out3)[:, :, 0], ref3, atol=1e-4, rtol=1e-4)


def test_fused_affine_silu_on_load():
    torch.manual_seed(5)
    B, D, H, W = 2, 2, 4, 4
    x = torch.randn(B, 16, D, H, W).half().float()
    a = torch.randn(B, 16)
    c = torch.randn(B, 16)
    w = (torch.randn(16, 16, 3, 3, 3) * 0.1).half().float()
    plan = TapGemm(w, None, device="cpu", n_tile=16)
    out, _ = emulate(plan, cl(x), coef0=(a, c))
    act = F.silu(x * a[:, :, None, None, None] + c[:, :, None, None, None]).half().float()
    ref = F.conv3d(act, w, padding=1)
    assert torch.allclose(uncl(out), ref, atol=2e-3, rtol=1e-3)


def test_column_strips_7x7_and_3x3(monkeypatch):
    """strip mode (p.strips > 1): every plane cut into column strips with real neighbour columns as halo -- forced here
    on small grids (the planner only picks it for wide planes whose kz-stacked plan does not fit otherwise), for the
    generic and the kz-stacked issue schemes, including a width that the strips do not divide."""
    torch.manual_seed(6)
    for (cin, cout, k, B, D, H, W, strips, ntile) in ((16, 64, 7, 1, 4, 5, 11, 2, None), (16, 16, 3, 2, 2, 4, 9, 3, 16),
                                                      (32, 64, 3, 1, 5, 6, 8, 2, None)):
        x = torch.randn(B, cin, D, H, W).half().float()
        w = (torch.randn(cout, cin, k, k, k) * 0.05).half().float()
        bias = torch.randn(cout)
        monkeypatch.setenv("WDNO_FORCE_STRIPS", str(strips))
        plan = TapGemm(w, bias, device="cpu", n_tile=ntile, kc=16)
        p = plan._plan(B, D, H, W)
        assert p.strips == strips and p.Wfull == W and p.W == (W + strips - 1) // strips and p.Wp == p.W + 2 * (k // 2)
        out, stats = emulate(plan, cl(x), want_stats=True, groups=2)
        ref = F.conv3d(x, w, bias, padding=k // 2)
        assert torch.allclose(uncl(out), ref, atol=3e-4, rtol=1e-4), (cin, cout, k)
        rs = ref.reshape(B, 2, -1)
        assert torch.allclose(stats[:, :, 0].float(), rs.sum(-1), atol=2e-2)
        monkeypatch.delenv("WDNO_FORCE_STRIPS")


def test_planner_picks_strips_for_the_super_resolution_stem():
    """82(->96 padded) -> 64 channels, 7x7x7 on 24x80x80 (config C4): without strips the kz-stacked plan does not fit
    227 KB and the layer falls back to unstacked N = 64 MMAs (measured 261 TFLOP/s); with strips it fits"""
    w = torch.zeros(64, 82, 7, 7, 7)
    plan = TapGemm(w, None, src_channels=(96,), device="cpu")
    p = plan._plan(1, 24, 80, 80)
    assert p.zstack == 1 and p.strips == 2 and p.W == 40 and p.Wp == 46 and p.Wfull == 80
    p40 = TapGemm(torch.zeros(64, 42, 7, 7, 7), None, src_channels=(48,), device="cpu")._plan(1, 24, 40, 40)
    assert p40.zstack == 1 and p40.strips == 1 and p40.Wp == 43


def test_batch_folding_of_2d_layers(monkeypatch):
    """2-D layers (D == 1): samples folded into depth planes so that ZT of them share each weight tile; GroupNorm
    coefficients and statistics stay per sample (p.fold)"""
    torch.manual_seed(7)
    monkeypatch.setenv("WDNO_FOLD", "force")
    B, H, W = 8, 8, 8
    x = torch.randn(B, 32, 1, H, W).half().float()
    a, c = torch.randn(B, 32), torch.randn(B, 32)
    w = (torch.randn(128, 32, 3, 3) * 0.1).half().float()
    bias = torch.randn(128)
    plan = TapGemm(w, bias, device="cpu")
    p = plan._plan(B, 1, H, W)
    assert p.fold == 1 and p.B * p.D == B and p.D == 4 and p.ZT in (4, 2, 1)
    out, stats = emulate(plan, cl(x), coef0=(a, c), want_stats=True, groups=4)
    act = F.silu(x * a[:, :, None, None, None] + c[:, :, None, None, None]).half().float()
    ref = F.conv2d(act[:, :, 0], w, bias, padding=1)
    assert out.shape == (B, 1, H, W, 128)
    assert torch.allclose(uncl(out)[:, :, 0], ref, atol=3e-3, rtol=1e-3)
    rs = ref.reshape(B, 4, -1)
    assert torch.allclose(stats[:, :, 0].float(), rs.sum(-1), atol=5e-2, rtol=1e-3)
    assert torch.allclose(stats[:, :, 1].float(), (rs ** 2).sum(-1), rtol=2e-3)
    # the planner's own choice: folds the weight-heavy low-resolution layer, leaves the 64x64 layer alone
    monkeypatch.setenv("WDNO_FOLD", "1")
    heavy = TapGemm(torch.zeros(1024, 1024, 3, 3), None, device="cpu")._plan(256, 1, 8, 8)
    assert heavy.fold == 1 and heavy.ZT == 4
    light = TapGemm(torch.zeros(128, 128, 3, 3), None, device="cpu")._plan(256, 1, 64, 64)
    print("light plan", light.fold, light.ZT, light.PT)
