"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64-capable) of the wavelet arithmetic WDNO calls.

The arithmetic lives in third-party packages that are NOT vendored under /root/reference and are not installable
here: pytorch_wavelets (unpinned git HEAD, env.sh:7-10), ptwt==0.1.6 (env.sh:11), PyWavelets (unpinned, env.sh:6).
This file restates their published algorithms for the configurations the
reference's call sites use:
    DWTForward/DWTInverse(J, mode in {'zero','periodization'}, wave)   eval_ddpm_burgers.py:134-136,188-194;
                                                                      inference_2d.py:178-180,244-246
    DWT1DForward/DWT1DInverse                                          test_util.py:186-187; inference_2d.py:43-46
    ptwt.wavedec3 / waverec3 (level=1, mode='zero')                    inference_2d.py:41,141,184,220,250
PARITY PIN: the reference ships no golden vectors for this path and the libraries cannot be executed here, so the
restatement is pinned by (i) the coefficient shapes the reference documents (wave_trans_2d.py:172,176,183;
data_burgers_1d.py:53), (ii) perfect-reconstruction identities, (iii) analytic known answers (constant field,
unit impulse = outer product of taps, linear ramp has zero bior2.4 detail), (iv) 3-D == three 1-D passes.
=> "parity unpinned" against the real libraries (stated in DESIGN.md).
"""
import numpy as np

_S2 = np.sqrt(2.0)

# pywt.Wavelet(name).dec_lo / dec_hi  
_DEC = {
    "bior1.3": (_S2 * np.array([-1, 1, 8, 8, 1, -1], dtype=np.float64) / 16.0,
                _S2 * np.array([0, 0, -1, 1, 0, 0], dtype=np.float64) / 2.0),
    "bior2.4": (_S2 * np.array([0, 3, -6, -16, 38, 90, 38, -16, -6, 3], dtype=np.float64) / 128.0,
                _S2 * np.array([0, 0, 0, 1, -2, 1, 0, 0, 0, 0], dtype=np.float64) / 4.0),
    "haar": (_S2 * np.array([1, 1], dtype=np.float64) / 2.0, _S2 * np.array([-1, 1], dtype=np.float64) / 2.0),
}


def filter_bank(name):
    """-> dec_lo, dec_hi, rec_lo, rec_hi (pywt convention: rec_lo[k] = (-1)^(k+1) dec_hi[k], rec_hi[k] = (-1)^k dec_lo[k])"""
    dec_lo, dec_hi = _DEC[name]
    k = np.arange(len(dec_lo))
    rec_lo = ((-1.0) ** (k + 1)) * dec_hi
    rec_hi = ((-1.0) ** k) * dec_lo
    return dec_lo.copy(), dec_hi.copy(), rec_lo, rec_hi


# ---------------------------------------------------------------- 1-D analysis / synthesis along one axis
def _corr_stride2(xp, h, nout):
    """out[i] = sum_k xp[2i + k] * h[k] along the last axis"""
    L = len(h)
    out = np.zeros(xp.shape[:-1] + (nout,), dtype=xp.dtype)
    for k in range(L):
        out += xp[..., k:k + 2 * nout:2][..., :nout] * h[k]
    return out


def afb1d(x, wave, mode, axis=-1):
    """pytorch_wavelets.lowlevel.afb1d semantics (Appendix A.1).  -> (lo, hi) along `axis`."""
    dec_lo, dec_hi, _, _ = filter_bank(wave)
    h0, h1 = dec_lo[::-1], dec_hi[::-1]
    L = len(h0)
    x = np.moveaxis(np.asarray(x), axis, -1)
    N = x.shape[-1]
    if mode == "zero":
        nout = (N + L - 1) // 2
        p = 2 * (nout - 1) - N + L
        if p % 2 == 1:
            x = np.concatenate([x, np.zeros(x.shape[:-1] + (1,), x.dtype)], axis=-1)
        pad = p // 2
        z = np.zeros(x.shape[:-1] + (pad,), x.dtype)
        xp = np.concatenate([z, x, z], axis=-1)
        lo, hi = _corr_stride2(xp, h0, nout), _corr_stride2(xp, h1, nout)
    elif mode == "periodization":
        if N % 2 == 1:
            x = np.concatenate([x, x[..., -1:]], axis=-1)
            N += 1
        x = np.roll(x, -L // 2, axis=-1)
        z = np.zeros(x.shape[:-1] + (L - 1,), x.dtype)
        xp = np.concatenate([z, x, z], axis=-1)
        nfull = (N + 2 * (L - 1) - L) // 2 + 1
        lo, hi = _corr_stride2(xp, h0, nfull), _corr_stride2(xp, h1, nfull)
        N2 = N // 2
        for a in (lo, hi):
            a[..., :L // 2] += a[..., N2:N2 + L // 2]
        lo, hi = lo[..., :N2].copy(), hi[..., :N2].copy()
    else:
        raise ValueError(mode)
    return np.moveaxis(lo, -1, axis), np.moveaxis(hi, -1, axis)


def _convT_stride2(c, g):
    """y[2i + k] += c[i] * g[k] along the last axis; full length 2n + L - 2"""
    n, L = c.shape[-1], len(g)
    y = np.zeros(c.shape[:-1] + (2 * n + L - 2,), dtype=c.dtype)
    for k in range(L):
        y[..., k:k + 2 * n:2] += c * g[k]
    return y


def sfb1d(lo, hi, wave, mode, axis=-1):
    """pytorch_wavelets.lowlevel.sfb1d semantics (Appendix A.2)."""
    _, _, g0, g1 = filter_bank(wave)
    L = len(g0)
    lo = np.moveaxis(np.asarray(lo), axis, -1)
    hi = np.moveaxis(np.asarray(hi), axis, -1)
    n = lo.shape[-1]
    y = _convT_stride2(lo, g0) + _convT_stride2(hi, g1)
    if mode == "zero":
        y = y[..., L - 2: L - 2 + 2 * n - L + 2]
    elif mode == "periodization":
        y[..., :L - 2] += y[..., 2 * n:2 * n + L - 2]
        y = y[..., :2 * n]
        y = np.roll(y, 1 - L // 2, axis=-1)
    else:
        raise ValueError(mode)
    return np.moveaxis(y, -1, axis)


# ---------------------------------------------------------------- pytorch_wavelets-style 2-D / 1-D transforms
def dwt2_forward(x, J, wave, mode):
    """x [B,C,H,W] -> (Yl [B,C,H',W'], [Yh_j [B,C,3,H_j,W_j]] finest first); sub-bands (LH, HL, HH) =
    (lo_W*hi_H, hi_W*lo_H, hi_W*hi_H): row (W) pass first, then column (H) pass."""
    ll = np.asarray(x)
    yh = []
    for _ in range(J):
        lo_w, hi_w = afb1d(ll, wave, mode, axis=-1)
        ll_, lh = afb1d(lo_w, wave, mode, axis=-2)   # lo_W: (lo_H, hi_H)
        hl, hh = afb1d(hi_w, wave, mode, axis=-2)    # hi_W: (lo_H, hi_H)
        yh.append(np.stack([lh, hl, hh], axis=2))
        ll = ll_
    return ll, yh


def dwt2_inverse(yl, yh, wave, mode):
    ll = np.asarray(yl)
    for h in yh[::-1]:
        h = np.asarray(h)
        if ll.shape[-2] > h.shape[-2]:
            ll = ll[..., :-1, :]
        if ll.shape[-1] > h.shape[-1]:
            ll = ll[..., :-1]
        lh, hl, hhh = h[:, :, 0], h[:, :, 1], h[:, :, 2]
        lo = sfb1d(ll, lh, wave, mode, axis=-2)
        hi = sfb1d(hl, hhh, wave, mode, axis=-2)
        ll = sfb1d(lo, hi, wave, mode, axis=-1)
    return ll


def dwt1_forward(x, J, wave, mode):
    """x [B,C,N] -> (lo, [hi_j] finest first)"""
    lo = np.asarray(x)
    his = []
    for _ in range(J):
        lo, hi = afb1d(lo, wave, mode, axis=-1)
        his.append(hi)
    return lo, his


def dwt1_inverse(lo, his, wave, mode):
    lo = np.asarray(lo)
    for hi in his[::-1]:
        hi = np.asarray(hi)
        if lo.shape[-1] > hi.shape[-1]:
            lo = lo[..., :-1]
        lo = sfb1d(lo, hi, wave, mode, axis=-1)
    return lo


# ---------------------------------------------------------------- ptwt 0.1.6 wavedec3 / waverec3 (level 1, mode 'zero')
KEYS3 = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")


def _ptwt_analysis_axis(x, wave, axis):
    """pad (2L-3)//2 each side (+1 trailing if odd length), correlate with reversed dec filters, stride 2."""
    dec_lo, dec_hi, _, _ = filter_bank(wave)
    h0, h1 = dec_lo[::-1], dec_hi[::-1]
    L = len(h0)
    x = np.moveaxis(x, axis, -1)
    N = x.shape[-1]
    padl = padr = (2 * L - 3) // 2
    if N % 2 == 1:
        padr += 1
    xp = np.concatenate([np.zeros(x.shape[:-1] + (padl,), x.dtype), x, np.zeros(x.shape[:-1] + (padr,), x.dtype)], -1)
    nout = (xp.shape[-1] - L) // 2 + 1
    lo, hi = _corr_stride2(xp, h0, nout), _corr_stride2(xp, h1, nout)
    return np.moveaxis(lo, -1, axis), np.moveaxis(hi, -1, axis)


def wavedec3(x, wave, level=1, mode="zero"):
    """x [B,D,H,W] -> [aaa, {aad..ddd}] (letters index (D,H,W); a=lo, d=hi)"""
    assert level == 1 and mode == "zero"
    x = np.asarray(x)
    out = {}
    for kd, xd in zip("ad", _ptwt_analysis_axis(x, wave, 1)):
        for kh, xh in zip("ad", _ptwt_analysis_axis(xd, wave, 2)):
            for kw, xw in zip("ad", _ptwt_analysis_axis(xh, wave, 3)):
                out[kd + kh + kw] = xw
    return [out.pop("aaa"), {k: out[k] for k in KEYS3}]


def _ptwt_synthesis_axis(lo, hi, wave, axis):
    _, _, g0, g1 = filter_bank(wave)
    L = len(g0)
    lo, hi = np.moveaxis(lo, axis, -1), np.moveaxis(hi, axis, -1)
    y = _convT_stride2(lo, g0) + _convT_stride2(hi, g1)
    crop = (2 * L - 3) // 2
    y = y[..., crop:y.shape[-1] - crop]
    return np.moveaxis(y, -1, axis)


def waverec3(coeffs, wave):
    aaa, d = coeffs
    b = dict(d)
    b["aaa"] = np.asarray(aaa)
    b = {k: np.asarray(v) for k, v in b.items()}
    xd = {}
    for kd in "ad":
        xh = {}
        for kh in "ad":
            xh[kh] = _ptwt_synthesis_axis(b[kd + kh + "a"], b[kd + kh + "d"], wave, 3)
        xd[kd] = _ptwt_synthesis_axis(xh["a"], xh["d"], wave, 2)
    return _ptwt_synthesis_axis(xd["a"], xd["d"], wave, 1)
