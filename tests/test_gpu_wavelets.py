"""GPU parity of the DWT/IDWT kernels against the CPU restatement (oracle/wavelets.py), plus size-independent
properties at the full BASELINE sizes (perfect reconstruction, linearity, adjointness, sub-band permutation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-6  # fp32 kernels (fma accumulation) vs float64 oracle, relative to the largest magnitude


def maxrel(a, b):
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a.detach().cpu().double() - b).abs().max() / (b.abs().max() + 1e-30))


def test_wavedec3_waverec3_vs_oracle_and_ragged_sizes():
    from oracle import wavelets as O
    from wdno_b200 import wavelets as W
    rng = np.random.default_rng(0)
    for shape in ((3, 32, 64, 64), (2, 16, 64, 64), (1, 8, 32, 32), (2, 7, 9, 11), (1, 2, 2, 2)):
        x = rng.standard_normal(shape)
        xc = torch.tensor(x, dtype=torch.float32, device="cuda")
        aaa, d = W.wavedec3(xc, "bior1.3", mode="zero", level=1)
        aaa_o, d_o = O.wavedec3(x, "bior1.3")
        assert list(d.keys()) == list(O.KEYS3)
        assert tuple(aaa.shape) == aaa_o.shape and maxrel(aaa, aaa_o) < TOL
        for k in O.KEYS3:
            assert maxrel(d[k], d_o[k]) < TOL, (shape, k)
        rec = W.waverec3([aaa, d], W.Wavelet("bior1.3"))
        rec_o = O.waverec3([aaa_o, d_o], "bior1.3")
        assert tuple(rec.shape) == rec_o.shape and maxrel(rec, rec_o) < TOL
        if all(s % 2 == 0 for s in shape[1:]):
            assert maxrel(rec, x) < 5e-6


def test_dwt2d_and_1d_vs_oracle():
    from oracle import wavelets as O
    from wdno_b200 import wavelets as W
    rng = np.random.default_rng(1)
    cases = [((3, 2, 81, 120), 1, "bior2.4", "periodization"), ((2, 1, 64, 64), 1, "bior1.3", "zero"),
             ((1, 2, 64, 96), 3, "bior2.4", "periodization"), ((2, 3, 17, 23), 2, "bior1.3", "zero"),
             ((1, 1, 162, 240), 2, "bior2.4", "periodization")]
    for shape, J, wave, mode in cases:
        x = rng.standard_normal(shape)
        xc = torch.tensor(x, dtype=torch.float32, device="cuda")
        yl, yh = W.DWTForward(J=J, wave=wave, mode=mode)(xc)
        yl_o, yh_o = O.dwt2_forward(x, J, wave, mode)
        assert tuple(yl.shape) == yl_o.shape and maxrel(yl, yl_o) < TOL
        for a, b in zip(yh, yh_o):
            assert tuple(a.shape) == b.shape and maxrel(a, b) < TOL
        rec = W.DWTInverse(wave=wave, mode=mode)((yl, yh))
        rec_o = O.dwt2_inverse(yl_o, yh_o, wave, mode)
        assert tuple(rec.shape) == rec_o.shape and maxrel(rec, rec_o) < TOL
    for shape, J, wave, mode in [((2, 2, 120), 1, "bior2.4", "periodization"), ((2, 1, 32), 1, "bior1.3", "zero"),
                                 ((1, 3, 81), 2, "bior2.4", "periodization")]:
        x = rng.standard_normal(shape)
        xc = torch.tensor(x, dtype=torch.float32, device="cuda")
        lo, hi = W.DWT1DForward(J=J, wave=wave, mode=mode)(xc)
        lo_o, hi_o = O.dwt1_forward(x, J, wave, mode)
        assert maxrel(lo, lo_o) < TOL and all(maxrel(a, b) < TOL for a, b in zip(hi, hi_o))
        rec = W.DWT1DInverse(wave=wave, mode=mode)((lo, hi))
        assert maxrel(rec, O.dwt1_inverse(lo_o, hi_o, wave, mode)) < TOL


def test_full_size_properties():
    """BASELINE sizes: smoke 16 x 5 fields of 32x64x64, Burgers 256 x [2,81,120]"""
    from wdno_b200 import wavelets as W
    torch.manual_seed(0)
    x = torch.randn(80, 32, 64, 64, device="cuda")
    y = torch.randn_like(x)
    cx = W.wavedec3(x, "bior1.3")
    assert float((W.waverec3(cx, "bior1.3") - x).abs().max()) < 1e-5          # round trip
    cy = W.wavedec3(y, "bior1.3")
    cz = W.wavedec3(2.0 * x - 3.0 * y, "bior1.3")                             # linearity
    assert float((cz[0] - (2.0 * cx[0] - 3.0 * cy[0])).abs().max()) < 2e-5
    assert all(float((cz[1][k] - (2.0 * cx[1][k] - 3.0 * cy[1][k])).abs().max()) < 2e-5 for k in cx[1])
    # sub-band permutation: a field that only varies along W has energy only in the '..d' / 'aaa' bands
    ramp = torch.sin(torch.arange(64, device="cuda") * 0.7).expand(1, 32, 64, 64).contiguous()
    _, d = W.wavedec3(ramp, "bior1.3")
    inner = lambda t: t[:, 4:-4, 4:-4, 4:-4].abs().max()
    assert float(inner(d["aad"])) > 1e-2
    assert all(float(inner(d[k])) < 1e-5 for k in ("ada", "add", "daa", "dad", "dda", "ddd"))
    u = torch.randn(256, 2, 81, 120, device="cuda")
    yl, yh = W.DWTForward(J=1, wave="bior2.4", mode="periodization")(u)
    assert yl.shape == (256, 2, 41, 60) and yh[0].shape == (256, 2, 3, 41, 60)
    rec = W.DWTInverse(wave="bior2.4", mode="periodization")((yl, yh))
    assert rec.shape == (256, 2, 82, 120)
    assert float((rec[:, :, :81] - u).abs().max()) < 1e-5


def test_adjoint_backward_matches_autograd_definition():
    """<IDWT(c), g> == <c, IDWT^T(g)> : the backward kernels are exact adjoints (gradient guidance path)."""
    from wdno_b200 import wavelets as W
    torch.manual_seed(1)
    aaa = torch.randn(2, 18, 34, 34, device="cuda", requires_grad=True)
    d = {k: torch.randn(2, 18, 34, 34, device="cuda", requires_grad=True) for k in W.KEYS3}
    rec = W.waverec3([aaa, d], "bior1.3")
    g = torch.randn_like(rec)
    (rec * g).sum().backward()
    # adjoint identity with the forward transform: IDWT^T = analysis with the synthesis filters
    lhs = float((rec.detach() * g).sum())
    rhs = float((aaa.detach() * aaa.grad).sum() + sum((d[k].detach() * d[k].grad).sum() for k in d))
    assert abs(lhs - rhs) < 1e-3 * abs(lhs)
    yl = torch.randn(2, 2, 41, 60, device="cuda", requires_grad=True)
    yh = torch.randn(2, 2, 3, 41, 60, device="cuda", requires_grad=True)
    rec = W.DWTInverse(wave="bior2.4", mode="periodization")((yl, [yh]))
    g = torch.randn_like(rec)
    (rec * g).sum().backward()
    lhs = float((rec.detach() * g).sum())
    rhs = float((yl.detach() * yl.grad).sum() + (yh.detach() * yh.grad).sum())
    assert abs(lhs - rhs) < 1e-3 * abs(lhs)
    # finite-difference check of the forward-transform backward (periodization, odd length 81)
    x = torch.randn(1, 1, 81, 12, device="cuda", requires_grad=True)
    yl, yh = W.DWTForward(J=1, wave="bior2.4", mode="periodization")(x)
    w1, w2 = torch.randn_like(yl), torch.randn_like(yh[0])
    ((yl * w1).sum() + (yh[0] * w2).sum()).backward()
    dx = torch.randn_like(x)
    with torch.no_grad():
        yl2, yh2 = W.DWTForward(J=1, wave="bior2.4", mode="periodization")(x + dx)
        fd = float(((yl2 - yl) * w1).sum() + ((yh2[0] - yh[0]) * w2).sum())
    assert abs(fd - float((x.grad * dx).sum())) < 1e-3 * max(1.0, abs(fd))


def test_fused_2d_kernels_bit_equal_to_separable_passes_and_strips(monkeypatch):
    """csrc/dwt2d.cu (one launch per level) against the three per-axis launches of csrc/dwt.cu: same formulas and summation
    order -> identical bits; images too tall for one shared-memory tile are cut into row strips (300 x 200, 162 x 240)."""
    from oracle import wavelets as O
    from wdno_b200 import wavelets as W
    rng = np.random.default_rng(7)
    cases = [((3, 2, 81, 120), "bior2.4", "periodization"), ((2, 1, 64, 64), "bior1.3", "zero"), ((1, 1, 300, 200), "bior2.4", "periodization"),
             ((1, 2, 162, 240), "bior2.4", "periodization"), ((2, 3, 17, 23), "bior1.3", "zero"), ((1, 1, 301, 130), "bior1.3", "zero"),
             ((2, 1, 21, 30), "bior2.4", "periodization"), ((1, 1, 10, 12), "bior2.4", "zero")]
    for shape, wave, mode in cases:
        x = rng.standard_normal(shape)
        xc = torch.tensor(x, dtype=torch.float32, device="cuda")
        res = {}
        for fused in (True, False):
            monkeypatch.setattr(W, "_FUSED2D", fused)
            yl, yh = W.DWTForward(J=1, wave=wave, mode=mode)(xc)
            rec = W.DWTInverse(wave=wave, mode=mode)((yl, yh))
            res[fused] = (yl, yh[0], rec, W.dwt2_packed(xc, wave, mode))
        for a, b in zip(res[True], res[False]):
            assert a.shape == b.shape and torch.equal(a, b), (shape, wave, mode)
        yl_o, yh_o = O.dwt2_forward(x, 1, wave, mode)
        assert maxrel(res[True][0], yl_o) < TOL and maxrel(res[True][1], yh_o[0]) < TOL
        assert maxrel(res[True][2], O.dwt2_inverse(yl_o, yh_o, wave, mode)) < TOL
    # gradients through the fused forward passes equal those through the separable ones (the backward is per-axis in both)
    grads = {}
    for fused in (True, False):
        monkeypatch.setattr(W, "_FUSED2D", fused)
        torch.manual_seed(3)
        x = torch.randn(2, 2, 81, 120, device="cuda", requires_grad=True)
        yl, yh = W.DWTForward(J=2, wave="bior2.4", mode="periodization")(x)
        rec = W.DWTInverse(wave="bior2.4", mode="periodization")((yl * 1.5, [h * 0.5 for h in yh]))
        (rec.square().sum() + yl.sum()).backward()
        grads[fused] = x.grad.clone()
    assert float((grads[True] - grads[False]).abs().max()) <= 1e-5 * float(grads[False].abs().max())


def test_fused_dwt2d_more_shapes_equal_to_separable_passes():
    """the fused 2-D kernels (extension-staged synthesis, tap-mask variants) against the per-axis passes on odd / tiny shapes"""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np, torch
from wdno_b200 import wavelets as W
rng = np.random.default_rng(11)
for shape, wave, mode in (((3, 2, 81, 120), "bior2.4", "periodization"), ((2, 1, 64, 64), "bior1.3", "zero"), ((1, 1, 300, 200), "bior2.4", "periodization"),
                          ((2, 3, 17, 23), "bior1.3", "zero"), ((2, 1, 21, 30), "bior2.4", "periodization"), ((1, 1, 10, 12), "bior2.4", "zero")):
    x = torch.tensor(rng.standard_normal(shape), dtype=torch.float32, device="cuda")
    res = {}
    for fused in (True, False):
        W._FUSED2D = fused
        yl, yh = W.DWTForward(J=1, wave=wave, mode=mode)(x)
        res[fused] = (yl, yh[0], W.DWTInverse(wave=wave, mode=mode)((yl, yh)))
    for a, b in zip(res[True], res[False]):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-6 * float(b.abs().max()), (shape, wave, mode)
print("v2 ok")
'''
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "v2 ok" in r.stdout, r.stdout + r.stderr
