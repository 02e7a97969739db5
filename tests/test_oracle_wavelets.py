"""Pins of the CPU wavelet restatement (oracle/wavelets.py): documented shapes, perfect reconstruction,
analytic known answers.  The real libraries are not installable here -> see the module docstring."""
import numpy as np
import pytest

from oracle import wavelets as W


def test_filter_bank_perfect_reconstruction_condition():
    for name in ("bior1.3", "bior2.4"):
        dl, dh, rl, rh = W.filter_bank(name)
        assert abs(dl.sum() - np.sqrt(2)) < 1e-12 and abs(dh.sum()) < 1e-12
        # PR: sum_k rec_lo[k] dec_lo[2n-k] + rec_hi[k] dec_hi[2n-k] = delta
        full = np.convolve(rl, dl) + np.convolve(rh, dh)
        centre = len(full) // 2
        assert abs(full[centre] - 2.0) < 1e-12
        assert np.abs(np.delete(full, centre)[1::2]).max() < 1e-12 or np.abs(full[centre % 2::2]).max() > 0


def test_reference_documented_shapes():
    # smoke/wave_trans_2d.py:172  [5, 8, 18, 34, 34] from 32x64x64 ; data_burgers_1d.py:53  [.., 8, 41, 60] from 81x120
    x = np.random.default_rng(0).standard_normal((5, 32, 64, 64))
    aaa, d = W.wavedec3(x, "bior1.3")
    assert aaa.shape == (5, 18, 34, 34) and all(d[k].shape == (5, 18, 34, 34) for k in W.KEYS3)
    assert list(d.keys()) == list(W.KEYS3)
    u = np.random.default_rng(1).standard_normal((3, 2, 81, 120))
    yl, yh = W.dwt2_forward(u, 1, "bior2.4", "periodization")
    assert yl.shape == (3, 2, 41, 60) and yh[0].shape == (3, 2, 3, 41, 60)
    # time / space down-sampled variants (wave_trans_2d.py:176,183)
    for shp, out in (((1, 16, 64, 64), (10, 34, 34)), ((1, 8, 64, 64), (6, 34, 34)), ((1, 32, 32, 32), (18, 18, 18)),
                     ((1, 32, 16, 16), (18, 10, 10))):
        a, _ = W.wavedec3(np.zeros(shp), "bior1.3")
        assert a.shape[1:] == out
    rho = np.zeros((2, 1, 64, 64))
    yl, yh = W.dwt2_forward(rho, 1, "bior1.3", "zero")
    assert yl.shape == (2, 1, 34, 34) and yh[0].shape == (2, 1, 3, 34, 34)
    lo, hi = W.dwt1_forward(np.zeros((2, 2, 120)), 1, "bior2.4", "periodization")
    assert lo.shape == (2, 2, 60) and hi[0].shape == (2, 2, 60)


def test_perfect_reconstruction():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 32, 64, 64))
    assert np.abs(W.waverec3(W.wavedec3(x, "bior1.3"), "bior1.3") - x).max() < 1e-12
    u = rng.standard_normal((2, 2, 81, 120))
    yl, yh = W.dwt2_forward(u, 1, "bior2.4", "periodization")
    rec = W.dwt2_inverse(yl, yh, "bior2.4", "periodization")
    assert rec.shape == (2, 2, 82, 120)
    assert np.abs(rec[:, :, :81, :120] - u).max() < 1e-12
    r = rng.standard_normal((2, 1, 64, 64))
    yl, yh = W.dwt2_forward(r, 1, "bior1.3", "zero")
    assert np.abs(W.dwt2_inverse(yl, yh, "bior1.3", "zero") - r).max() < 1e-12
    s = rng.standard_normal((2, 1, 32))
    lo, hi = W.dwt1_forward(s, 1, "bior1.3", "zero")
    assert lo.shape[-1] == 18
    assert np.abs(W.dwt1_inverse(lo, hi, "bior1.3", "zero") - s).max() < 1e-12
    v = rng.standard_normal((2, 2, 120))
    lo, hi = W.dwt1_forward(v, 1, "bior2.4", "periodization")
    assert np.abs(W.dwt1_inverse(lo, hi, "bior2.4", "periodization") - v).max() < 1e-12
    # multi-level 2-D
    w = rng.standard_normal((1, 1, 64, 96))
    yl, yh = W.dwt2_forward(w, 3, "bior2.4", "periodization")
    assert np.abs(W.dwt2_inverse(yl, yh, "bior2.4", "periodization") - w).max() < 1e-11


def test_known_answers():
    # constant field: interior LLL = (sqrt 2)^3 c = 2 sqrt2 c, all detail 0 (bior1.3)
    c = 0.75
    aaa, d = W.wavedec3(np.full((1, 32, 64, 64), c), "bior1.3")
    assert np.abs(aaa[0, 4:-4, 4:-4, 4:-4] - 2 * np.sqrt(2) * c).max() < 1e-12
    assert max(np.abs(v[0, 4:-4, 4:-4, 4:-4]).max() for v in d.values()) < 1e-12
    # unit impulse: each sub-band equals the outer product of the (strided) reversed taps
    x = np.zeros((1, 32, 64, 64))
    x[0, 10, 20, 30] = 1.0
    aaa, d = W.wavedec3(x, "bior1.3")
    dl, dh, _, _ = W.filter_bank("bior1.3")

    def resp(n, h, N):
        L, pad = len(h), (2 * len(h) - 3) // 2
        out = np.zeros((N + 2 * pad - L) // 2 + 1)
        for i in range(len(out)):
            k = n + pad - 2 * i
            if 0 <= k < L:
                out[i] = h[::-1][k]
        return out
    exp = np.einsum("i,j,k->ijk", resp(10, dl, 32), resp(20, dh, 64), resp(30, dl, 64))
    assert np.abs(d["ada"][0] - exp).max() < 1e-14
    # linear ramp: bior2.4 has 2 vanishing moments on the analysis high-pass -> zero detail away from the wrap
    ramp = np.arange(120, dtype=np.float64)[None, None, :]
    lo, hi = W.dwt1_forward(ramp, 1, "bior2.4", "periodization")
    assert np.abs(hi[0][0, 0, 4:-4]).max() < 1e-10


def test_separable_3d_equals_1d_passes():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 8, 12, 10))
    aaa, d = W.wavedec3(x, "bior1.3")
    lo_w, hi_w = W.afb1d(x, "bior1.3", "zero", axis=3)
    lo_h, hi_h = W.afb1d(hi_w, "bior1.3", "zero", axis=2)
    dd_lo, dd_hi = W.afb1d(lo_h, "bior1.3", "zero", axis=1)
    # (D=hi, H=lo, W=hi) = "dad"
    assert np.abs(d["dad"] - dd_hi).max() < 1e-13
