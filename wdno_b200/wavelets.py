"""DWT / IDWT with the call signatures of the third-party libraries WDNO uses, on the CUDA kernels of csrc/dwt.cu.

    DWTForward(J, mode, wave)(x) -> (Yl, [Yh...])       pytorch_wavelets       (inference_2d.py:178-180,244-246;
    DWTInverse(mode, wave)((Yl, Yh)) -> x                                       eval_ddpm_burgers.py:134-136,188-194)
    DWT1DForward / DWT1DInverse                          pytorch_wavelets       (test_util.py:186-187; inference_2d.py:43-46)
    wavedec3(x, wavelet, mode='zero', level=1) -> [Yl, {aad..ddd}]   ptwt 0.1.6 (inference_2d.py:41,141,184,250)
    waverec3([Yl, {...}], wavelet) -> x                                        (inference_2d.py:141,220)
    Wavelet(name)                                        pywt.Wavelet stand-in (filter banks of SURVEY.md row a21)

Every transform is differentiable (torch.autograd.Function whose backward is the adjoint pass on the same kernels),
which is what the reference's gradient guidance needs (inference_2d.py:30-66, eval_ddpm_burgers.py:108-147).
fp32 CUDA tensors only; there is no CPU path.
"""
import ctypes as C
import math
import os

import torch
from torch import nn

from . import _lib

_S2 = math.sqrt(2.0)
_DEC = {
    "bior1.3": ([_S2 * v / 16.0 for v in (-1, 1, 8, 8, 1, -1)], [_S2 * v / 2.0 for v in (0, 0, -1, 1, 0, 0)]),
    "bior2.4": ([_S2 * v / 128.0 for v in (0, 3, -6, -16, 38, 90, 38, -16, -6, 3)],
                [_S2 * v / 4.0 for v in (0, 0, 0, 1, -2, 1, 0, 0, 0, 0)]),
    "haar": ([_S2 / 2.0, _S2 / 2.0], [-_S2 / 2.0, _S2 / 2.0]),
}


class Wavelet:
    """pywt.Wavelet stand-in: .name, .dec_lo, .dec_hi, .rec_lo, .rec_hi, .dec_len"""

    def __init__(self, name):
        if isinstance(name, Wavelet):
            name = name.name
        if hasattr(name, "name") and not isinstance(name, str):
            name = name.name
        if name not in _DEC:
            raise ValueError(f"wavelet {name!r} not built (have {sorted(_DEC)})")
        self.name = name
        self.dec_lo, self.dec_hi = list(_DEC[name][0]), list(_DEC[name][1])
        self.rec_lo = [((-1.0) ** (k + 1)) * v for k, v in enumerate(self.dec_hi)]
        self.rec_hi = [((-1.0) ** k) * v for k, v in enumerate(self.dec_lo)]
        self.dec_len = self.rec_len = len(self.dec_lo)

    @property
    def filter_bank(self):
        return self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi


def _wave(w):
    return w if isinstance(w, Wavelet) else Wavelet(w)


def _farr(v):
    return (C.c_float * len(v))(*v)


def _check(x):
    if not (torch.is_tensor(x) and x.is_cuda):
        raise RuntimeError("wdno_b200.wavelets: CUDA tensors only (no CPU path)")


# ---------------------------------------------------------------- raw axis passes
def _mode_id(mode):
    if mode in ("zero", "constant"):
        return 0
    if mode in ("periodization", "per"):
        return 1
    raise ValueError(f"padding mode {mode!r} not built (zero / periodization)")


def _analysis_raw(x, axis, t_lo, t_hi, off, periodic, nout, lo=None, hi=None):
    """x contiguous fp32; returns (lo, hi) (or writes into given views that are contiguous beyond `axis`)."""
    x = x.contiguous()
    axis = axis % x.dim()
    N = x.shape[axis]
    outer = int(math.prod(x.shape[:axis])) if axis > 0 else 1
    inner = int(math.prod(x.shape[axis + 1:])) if axis < x.dim() - 1 else 1
    oshape = list(x.shape)
    oshape[axis] = nout
    if lo is None:
        lo = torch.empty(oshape, dtype=torch.float32, device=x.device)
    if hi is None:
        hi = torch.empty(oshape, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().wdno_dwt_analysis_axis(
        x.data_ptr(), lo.data_ptr(), hi.data_ptr(), outer, N, inner, nout, N * inner, _ostride(lo, axis, nout, inner),
        _ostride(hi, axis, nout, inner), _farr(t_lo), _farr(t_hi), len(t_lo), off, periodic, _lib.current_stream_ptr()),
        "dwt_analysis_axis")
    return lo, hi


def _ostride(t, axis, n, inner):
    """outer stride (elements) of a tensor/view whose dims from `axis` on are contiguous and dims before it are
    uniformly strided (plain tensors, or a band slice of a stacked tensor)"""
    if t.dim() == 0 or axis == 0:
        return n * inner
    assert t.stride(axis) == inner, "sub-band view must be contiguous from the transform axis on"
    return t.stride(axis - 1)


def _synthesis_raw(lo, hi, axis, t_lo, t_hi, off, periodic, nout):
    axis = axis % lo.dim()
    n = lo.shape[axis]
    outer = int(math.prod(lo.shape[:axis])) if axis > 0 else 1
    inner = int(math.prod(lo.shape[axis + 1:])) if axis < lo.dim() - 1 else 1
    oshape = list(lo.shape)
    oshape[axis] = nout
    y = torch.empty(oshape, dtype=torch.float32, device=lo.device)
    _lib.check(_lib.lib().wdno_dwt_synthesis_axis(
        lo.data_ptr(), hi.data_ptr(), y.data_ptr(), outer, n, inner, nout, _ostride(lo, axis, n, inner),
        _ostride(hi, axis, n, inner), nout * inner, _farr(t_lo), _farr(t_hi), len(t_lo), off, periodic,
        _lib.current_stream_ptr()), "dwt_synthesis_axis")
    return y


def _uniform_outer(t, axis):
    """True if dims before `axis` collapse to one uniformly strided outer dim and dims from axis on are contiguous"""
    if not t.is_cuda:
        return False
    exp = 1
    for d in range(t.dim() - 1, axis - 1, -1):
        if t.shape[d] != 1 and t.stride(d) != exp:
            return False
        exp *= t.shape[d]
    st = None
    for d in range(axis - 1, -1, -1):
        if st is None:
            st = t.stride(d)
            run = st * t.shape[d]
        else:
            if t.shape[d] != 1 and t.stride(d) != run:
                return False
            run = t.stride(d) * t.shape[d]
    return True


def _prep(t, axis):
    t = t.to(torch.float32)
    return t if _uniform_outer(t, axis % t.dim()) else t.contiguous()


def _geom(N, L, mode):
    """-> (nout, off, periodic) of the analysis along an axis of length N"""
    if _mode_id(mode) == 0:
        return (N + L - 1) // 2, L - 2, 0
    return (N + (N & 1)) // 2, L // 2 - 1, 1


class _Analysis(torch.autograd.Function):
    """(lo, hi) = afb1d(x) along `axis` (Appendix A.1 / A.3)"""

    @staticmethod
    def forward(ctx, x, axis, wname, mode):
        w = Wavelet(wname)
        L = w.dec_len
        x = x.to(torch.float32).contiguous()
        N = x.shape[axis]
        nout, off, per = _geom(N, L, mode)
        ctx.cfg = (axis, wname, mode, N)
        return _analysis_raw(x, axis, w.dec_lo[::-1], w.dec_hi[::-1], off, per, nout)

    @staticmethod
    def backward(ctx, glo, ghi):
        axis, wname, mode, N = ctx.cfg
        return _analysis_adjoint(glo, ghi, axis, wname, mode, N), None, None, None


def _analysis_adjoint(glo, ghi, axis, wname, mode, N):
    """gradient of (lo, hi) = afb1d(x) w.r.t. x: the synthesis kernel with the analysis taps (+ the odd-N fold of 'per')"""
    w = Wavelet(wname)
    L = w.dec_len
    _, off, per = _geom(N, L, mode)
    glo, ghi = _prep(glo, axis), _prep(ghi, axis)
    if per:
        Np = N + (N & 1)
        g = _synthesis_raw(glo, ghi, axis, w.dec_lo[::-1], w.dec_hi[::-1], off, 1, Np)
        if Np != N:  # the repeated edge sample folds back onto the last real sample
            idx = [slice(None)] * g.dim()
            idx[axis] = slice(0, N)
            gx = g[tuple(idx)].clone()
            last, ext = list(idx), list(idx)
            last[axis], ext[axis] = slice(N - 1, N), slice(N, N + 1)
            gx[tuple(last)] += g[tuple(ext)]
            g = gx
    else:
        g = _synthesis_raw(glo, ghi, axis, w.dec_lo[::-1], w.dec_hi[::-1], off, 0, N)
    return g


class _Synthesis(torch.autograd.Function):
    """y = sfb1d(lo, hi) along `axis` (Appendix A.2 / A.3)"""

    @staticmethod
    def forward(ctx, lo, hi, axis, wname, mode):
        w = Wavelet(wname)
        L = w.dec_len
        lo, hi = _prep(lo, axis), _prep(hi, axis)
        n = lo.shape[axis]
        per = _mode_id(mode)
        nout = 2 * n if per else 2 * n - L + 2
        off = L // 2 - 1 if per else L - 2
        ctx.cfg = (axis, wname, mode, n, off, per)
        return _synthesis_raw(lo, hi, axis, w.rec_lo, w.rec_hi, off, per, nout)

    @staticmethod
    def backward(ctx, gy):
        axis, wname, mode, n, off, per = ctx.cfg
        w = Wavelet(wname)
        gy = gy.to(torch.float32).contiguous()
        glo, ghi = _analysis_raw(gy, axis, w.rec_lo, w.rec_hi, off, per, n)
        return glo, ghi, None, None, None


def afb1d(x, wave, mode, axis=-1):
    _check(x)
    return _Analysis.apply(x, axis % x.dim(), _wave(wave).name, mode)


def sfb1d(lo, hi, wave, mode, axis=-1):
    _check(lo)
    return _Synthesis.apply(lo, hi, axis % lo.dim(), _wave(wave).name, mode)


# ---------------------------------------------------------------- fused one-level 2-D transforms (csrc/dwt2d.cu)
_FUSED2D = os.environ.get("WDNO_DWT2D_FUSED", "1") != "0"   # 0: three per-axis launches per level (csrc/dwt.cu)


def _img_stride(t):
    """element stride between consecutive images of t [..., h, w] when the leading dims collapse to one uniformly strided
    image index and the planes are contiguous; None otherwise"""
    h, w = t.shape[-2], t.shape[-1]
    if (w != 1 and t.stride(-1) != 1) or (h != 1 and t.stride(-2) != w):
        return None
    st = run = None
    for d in range(t.dim() - 3, -1, -1):
        if t.shape[d] == 1:
            continue
        if st is None:
            st, run = t.stride(d), t.stride(d) * t.shape[d]
        elif t.stride(d) != run:
            return None
        else:
            run = t.stride(d) * t.shape[d]
    return h * w if st is None else st


def _planes(t):
    """-> (tensor usable by the 2-D kernels, image stride)"""
    st = _img_stride(t) if t.dtype == torch.float32 else None
    if st is None:
        t = t.to(torch.float32).contiguous()
        st = t.shape[-2] * t.shape[-1]
    return t, st


def _fused2d_ok(L, H, W, nh, nw, per, t):
    n_img = t.numel() // max(1, t.shape[-2] * t.shape[-1])
    return (_FUSED2D and t.is_cuda and t.dim() >= 2 and 1 <= n_img <= 65535
            and bool(_lib.lib().wdno_dwt2d_supported(L, H, W, nh, nw, per)))


def _ana2d_raw(x, bands, t_lo, t_hi, geo_h, geo_w):
    """x [..., H, W] contiguous fp32 -> the four given band views [..., nh, nw] (ll, lh, hl, hh), one launch"""
    H, W = x.shape[-2], x.shape[-1]
    n_img = x.numel() // (H * W)
    strides = []
    for b in bands:
        st = _img_stride(b)
        assert st is not None, "sub-band views must have contiguous planes"
        strides.append(st)
    _lib.check(_lib.lib().wdno_dwt2d_analysis(
        x.data_ptr(), (C.c_void_p * 4)(*[b.data_ptr() for b in bands]), (C.c_int64 * 4)(*strides), n_img, H, W, geo_h[0], geo_w[0],
        _farr(t_lo), _farr(t_hi), len(t_lo), geo_h[1], geo_w[1], geo_h[2], _lib.current_stream_ptr()), "dwt2d_analysis")


def _syn2d_raw(bands, t_lo, t_hi, off, per, HW):
    """(ll, lh, hl, hh) [..., nh, nw] -> y [..., H, W], one launch"""
    prepared = [_planes(b) for b in bands]
    nh, nw = bands[0].shape[-2], bands[0].shape[-1]
    lead = tuple(bands[0].shape[:-2])
    y = torch.empty(lead + tuple(HW), dtype=torch.float32, device=bands[0].device)
    n_img = y.numel() // (HW[0] * HW[1])
    _lib.check(_lib.lib().wdno_dwt2d_synthesis(
        (C.c_void_p * 4)(*[b.data_ptr() for b, _ in prepared]), (C.c_int64 * 4)(*[st for _, st in prepared]), y.data_ptr(), n_img,
        nh, nw, HW[0], HW[1], _farr(t_lo), _farr(t_hi), len(t_lo), off, off, per, _lib.current_stream_ptr()), "dwt2d_synthesis")
    return y


class _Analysis2D(torch.autograd.Function):
    """(ll [..., nh, nw], yh [..., 3, nh, nw]) = one DWTForward level of x [..., H, W]; backward = the per-axis adjoints"""

    @staticmethod
    def forward(ctx, x, wname, mode):
        w = Wavelet(wname)
        L = w.dec_len
        x = x.to(torch.float32).contiguous()
        H, W = x.shape[-2], x.shape[-1]
        gh, gw = _geom(H, L, mode), _geom(W, L, mode)
        lead = tuple(x.shape[:-2])
        ll = torch.empty(lead + (gh[0], gw[0]), dtype=torch.float32, device=x.device)
        yh = torch.empty(lead + (3, gh[0], gw[0]), dtype=torch.float32, device=x.device)
        _ana2d_raw(x, [ll, yh.select(-3, 0), yh.select(-3, 1), yh.select(-3, 2)], w.dec_lo[::-1], w.dec_hi[::-1], gh, gw)
        ctx.cfg = (wname, mode, H, W)
        return ll, yh

    @staticmethod
    def backward(ctx, gll, gyh):
        wname, mode, H, W = ctx.cfg
        ax_h, ax_w = gll.dim() - 2, gll.dim() - 1
        g_lo_w = _analysis_adjoint(gll, gyh.select(-3, 0), ax_h, wname, mode, H)
        g_hi_w = _analysis_adjoint(gyh.select(-3, 1), gyh.select(-3, 2), ax_h, wname, mode, H)
        return _analysis_adjoint(g_lo_w, g_hi_w, ax_w, wname, mode, W), None, None


class _Synthesis2D(torch.autograd.Function):
    """y = one DWTInverse level of (ll, lh, hl, hh); backward = the per-axis analysis passes with the reconstruction taps"""

    @staticmethod
    def forward(ctx, ll, lh, hl, hh, wname, mode):
        w = Wavelet(wname)
        L = w.dec_len
        per = _mode_id(mode)
        nh, nw = ll.shape[-2], ll.shape[-1]
        off = L // 2 - 1 if per else L - 2
        HW = (2 * nh, 2 * nw) if per else (2 * nh - L + 2, 2 * nw - L + 2)
        ctx.cfg = (wname, nh, nw, off, per)
        return _syn2d_raw([ll, lh, hl, hh], w.rec_lo, w.rec_hi, off, per, HW)

    @staticmethod
    def backward(ctx, gy):
        wname, nh, nw, off, per = ctx.cfg
        w = Wavelet(wname)
        gy = gy.to(torch.float32).contiguous()
        ax_h, ax_w = gy.dim() - 2, gy.dim() - 1
        g_lo, g_hi = _analysis_raw(gy, ax_w, w.rec_lo, w.rec_hi, off, per, nw)
        g_ll, g_lh = _analysis_raw(g_lo, ax_h, w.rec_lo, w.rec_hi, off, per, nh)
        g_hl, g_hh = _analysis_raw(g_hi, ax_h, w.rec_lo, w.rec_hi, off, per, nh)
        return g_ll, g_lh, g_hl, g_hh, None, None


def _dwt2_level(x, w, mode):
    """one analysis level of x [..., H, W] -> (ll, yh [..., 3, nh, nw])"""
    L = w.dec_len
    H, W = x.shape[-2], x.shape[-1]
    gh, gw = _geom(H, L, mode), _geom(W, L, mode)
    if _fused2d_ok(L, H, W, gh[0], gw[0], gh[2], x):
        return _Analysis2D.apply(x, w.name, mode)
    lo_w, hi_w = afb1d(x, w, mode, axis=-1)
    ll, lh = afb1d(lo_w, w, mode, axis=-2)
    hl, hh = afb1d(hi_w, w, mode, axis=-2)
    return ll, torch.stack((lh, hl, hh), dim=-3)


def _idwt2_level(ll, h, w, mode):
    """one synthesis level: ll [..., nh, nw], h [..., 3, nh, nw] -> [..., H, W]"""
    L = w.dec_len
    per = _mode_id(mode)
    nh, nw = ll.shape[-2], ll.shape[-1]
    HW = (2 * nh, 2 * nw) if per else (2 * nh - L + 2, 2 * nw - L + 2)
    if HW[0] >= 1 and HW[1] >= 1 and _fused2d_ok(L, HW[0], HW[1], nh, nw, per, ll):
        return _Synthesis2D.apply(ll, h.select(-3, 0), h.select(-3, 1), h.select(-3, 2), w.name, mode)
    lo = sfb1d(ll, h.select(-3, 0), w, mode, axis=-2)
    hi = sfb1d(h.select(-3, 1), h.select(-3, 2), w, mode, axis=-2)
    return sfb1d(lo, hi, w, mode, axis=-1)


# ---------------------------------------------------------------- pytorch_wavelets-style modules
class DWTForward(nn.Module):
    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        self.J, self.wave, self.mode = J, _wave(wave), mode

    def forward(self, x):
        """x [B,C,H,W] -> (Yl [B,C,H',W'], [Yh_j [B,C,3,H_j,W_j]], finest first); bands (LH, HL, HH)"""
        _check(x)
        ll, yh = x, []
        for _ in range(self.J):
            ll, h = _dwt2_level(ll, self.wave, self.mode)
            yh.append(h)
        return ll, yh


class DWTInverse(nn.Module):
    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        self.wave, self.mode = _wave(wave), mode

    def forward(self, coeffs):
        yl, yh = coeffs
        _check(yl)
        ll = yl
        for h in yh[::-1]:
            if h is None:
                h = torch.zeros(ll.shape[0], ll.shape[1], 3, ll.shape[-2], ll.shape[-1], device=ll.device)
            if ll.shape[-2] > h.shape[-2]:
                ll = ll[..., :-1, :]
            if ll.shape[-1] > h.shape[-1]:
                ll = ll[..., :-1]
            ll = _idwt2_level(ll, h, self.wave, self.mode)
        return ll


class DWT1DForward(nn.Module):
    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        self.J, self.wave, self.mode = J, _wave(wave), mode

    def forward(self, x):
        """x [B,C,N] -> (lo, [hi_j], finest first)"""
        _check(x)
        lo, his = x, []
        for _ in range(self.J):
            lo, hi = afb1d(lo, self.wave, self.mode, axis=-1)
            his.append(hi)
        return lo, his


class DWT1DInverse(nn.Module):
    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        self.wave, self.mode = _wave(wave), mode

    def forward(self, coeffs):
        lo, his = coeffs
        _check(lo)
        for hi in his[::-1]:
            if hi is None:
                hi = torch.zeros_like(lo)
            if lo.shape[-1] > hi.shape[-1]:
                lo = lo[..., :-1]
            lo = sfb1d(lo, hi, self.wave, self.mode, axis=-1)
        return lo


# ---------------------------------------------------------------- ptwt-style 3-D transform
KEYS3 = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")
_FUSED3D = os.environ.get("WDNO_DWT3D_FUSED", "1") != "0"   # csrc/dwt3d.cu: one launch per 3-D transform / adjoint


def _band_ptrs(bands):
    return (C.c_void_p * 8)(*[b.data_ptr() for b in bands])


def _ana3d_raw(x, t_lo, t_hi, off, nout):
    """x [B,Nd,Nh,Nw] contiguous fp32 -> stacked bands [8,B,nd,nh,nw] (band = 4 d + 2 h + w, 0 = low-pass)"""
    B, Nd, Nh, Nw = x.shape
    nd, nh, nw = nout
    out = torch.empty((8, B, nd, nh, nw), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().wdno_dwt3d_analysis(x.data_ptr(), _band_ptrs([out[i] for i in range(8)]), nd * nh * nw, B, Nd, Nh, Nw,
                                              nd, nh, nw, _farr(t_lo), _farr(t_hi), len(t_lo), off, _lib.current_stream_ptr()),
               "dwt3d_analysis")
    return out


def _syn3d_raw(bands, t_lo, t_hi, off, nout):
    """bands: 8 fp32 tensors [B,nd,nh,nw] (contiguous planes, one common batch stride) -> y [B,Nd,Nh,Nw]"""
    B, nd, nh, nw = bands[0].shape
    Nd, Nh, Nw = nout
    ok = all(b.shape == bands[0].shape and b.stride(3) == 1 and b.stride(2) == nw and b.stride(1) == nh * nw
             and b.stride(0) == bands[0].stride(0) for b in bands)
    if not ok:
        bands = [b.contiguous() for b in bands]
    y = torch.empty((B, Nd, Nh, Nw), dtype=torch.float32, device=bands[0].device)
    _lib.check(_lib.lib().wdno_dwt3d_synthesis(_band_ptrs(bands), bands[0].stride(0) if B > 1 else nd * nh * nw, y.data_ptr(), B,
                                               nd, nh, nw, Nd, Nh, Nw, _farr(t_lo), _farr(t_hi), len(t_lo), off,
                                               _lib.current_stream_ptr()), "dwt3d_synthesis")
    return y


def _fused3d_ok(L, nw, Nw, *tensors):
    return (_FUSED3D and all(t.dim() == 4 and t.is_cuda for t in tensors) and tensors[0].shape[0] <= 65535
            and bool(_lib.lib().wdno_dwt3d_supported(L, nw, Nw)))


class _Analysis3D(torch.autograd.Function):
    """stacked bands [8,B,nd,nh,nw] = wavedec3(x) ('zero'); backward = the fused synthesis with the same filters"""

    @staticmethod
    def forward(ctx, x, wname):
        w = Wavelet(wname)
        L = w.dec_len
        x = x.to(torch.float32).contiguous()
        geo = [_geom(n, L, "zero") for n in x.shape[1:]]
        ctx.cfg = (wname, tuple(x.shape[1:]), geo[0][1])
        return _ana3d_raw(x, w.dec_lo[::-1], w.dec_hi[::-1], geo[0][1], tuple(g[0] for g in geo))

    @staticmethod
    def backward(ctx, g):
        wname, N3, off = ctx.cfg
        w = Wavelet(wname)
        g = g.to(torch.float32).contiguous()
        return _syn3d_raw([g[i] for i in range(8)], w.dec_lo[::-1], w.dec_hi[::-1], off, N3), None


class _Synthesis3D(torch.autograd.Function):
    """y = waverec3(8 bands) ('zero'); backward = the fused analysis with the reconstruction filters"""

    @staticmethod
    def forward(ctx, wname, *bands):
        w = Wavelet(wname)
        L = w.dec_len
        bands = [b.to(torch.float32) for b in bands]
        n3 = tuple(bands[0].shape[1:])
        ctx.cfg = (wname, n3, L - 2)
        return _syn3d_raw(bands, w.rec_lo, w.rec_hi, L - 2, tuple(2 * n - L + 2 for n in n3))

    @staticmethod
    def backward(ctx, gy):
        wname, n3, off = ctx.cfg
        w = Wavelet(wname)
        g = _ana3d_raw(gy.to(torch.float32).contiguous(), w.rec_lo, w.rec_hi, off, n3)
        return (None,) + tuple(g[i] for i in range(8))



def wavedec3(data, wavelet, *, mode="zero", level=1):
    """data [B,D,H,W] -> [aaa, {aad, ada, add, daa, dad, dda, ddd}] ; letters index (D,H,W), a = low-pass"""
    _check(data)
    if level != 1:
        raise NotImplementedError("WDNO only uses level=1 (inference_2d.py:41,184,250)")
    if _mode_id(mode) != 0:
        raise NotImplementedError("WDNO only uses mode='zero' for the 3-D transform")
    w = _wave(wavelet)
    if _fused3d_ok(w.dec_len, _geom(data.shape[3], w.dec_len, "zero")[0], data.shape[3], data):
        bands = _Analysis3D.apply(data, w.name)
        return [bands[0], {k: bands[i + 1] for i, k in enumerate(KEYS3)}]
    out = {}
    for kd, xd in zip("ad", afb1d(data, w, "zero", axis=1)):
        for kh, xh in zip("ad", afb1d(xd, w, "zero", axis=2)):
            for kw, xw in zip("ad", afb1d(xh, w, "zero", axis=3)):
                out[kd + kh + kw] = xw
    return [out["aaa"], {k: out[k] for k in KEYS3}]


def waverec3(coeffs, wavelet):
    w = _wave(wavelet)
    aaa, d = coeffs[0], coeffs[1]
    _check(aaa)
    b = dict(d)
    b["aaa"] = aaa
    L = w.dec_len
    if all(k in b for k in KEYS3) and _fused3d_ok(L, aaa.shape[-1], 2 * aaa.shape[-1] - L + 2, *[b[k] for k in ("aaa",) + KEYS3]) \
            and all(b[k].shape == aaa.shape for k in KEYS3):
        return _Synthesis3D.apply(w.name, *[b[k] for k in ("aaa",) + KEYS3])
    xd = {}
    for kd in "ad":
        xh = {}
        for kh in "ad":
            xh[kh] = sfb1d(b[kd + kh + "a"], b[kd + kh + "d"], w, "zero", axis=3)
        xd[kd] = sfb1d(xh["a"], xh["d"], w, "zero", axis=2)
    return sfb1d(xd["a"], xd["d"], w, "zero", axis=1)


def waverec3_adjoint(gy, wavelet, coef_shape):
    """adjoint of `waverec3` without autograd: gy [B,D,H,W] (gradient w.r.t. the reconstructed fields) ->
    [g_aaa, {g_aad, ..., g_ddd}] each [B, *coef_shape] = the analysis kernel with the reconstruction filters (what
    `_Synthesis3D.backward` runs).  For objectives whose field gradient is known in closed form."""
    _check(gy)
    w = _wave(wavelet)
    L = w.dec_len
    gy = gy.detach().to(torch.float32).contiguous()
    n3 = tuple(int(v) for v in coef_shape[-3:])
    if _fused3d_ok(L, n3[2], gy.shape[3], gy):
        bands = _ana3d_raw(gy, w.rec_lo, w.rec_hi, L - 2, n3)
        return [bands[0], {k: bands[i + 1] for i, k in enumerate(KEYS3)}]
    out = {}
    for kd, xd in zip("ad", _analysis_raw(gy, 1, w.rec_lo, w.rec_hi, L - 2, 0, n3[0])):
        for kh, xh in zip("ad", _analysis_raw(xd, 2, w.rec_lo, w.rec_hi, L - 2, 0, n3[1])):
            for kw, xw in zip("ad", _analysis_raw(xh, 3, w.rec_lo, w.rec_hi, L - 2, 0, n3[2])):
                out[kd + kh + kw] = xw
    return [out["aaa"], {k: out[k] for k in KEYS3}]


# ---------------------------------------------------------------- packed forms for the offline coefficient builders
def wavedec3_packed(data, wavelet, *, mode="zero"):
    """data [B,D,H,W] -> [B,8,nd,nh,nw], bit-identical to smoke `coef_to_tensor(ptwt.wavedec3(data, ...))`
    (wave_trans_2d.py:55-58,129-130): the fused kernel writes the eight sub-bands straight into their slots of the
    packed tensor (band pointers = out[:, i], batch stride 8 nd nh nw), so the stack/cat pass never runs.  No autograd."""
    _check(data)
    if _mode_id(mode) != 0:
        raise NotImplementedError("WDNO only uses mode='zero' for the 3-D transform")
    w = _wave(wavelet)
    L = w.dec_len
    x = data.detach().to(torch.float32).contiguous()
    geo = [_geom(n, L, "zero") for n in x.shape[1:]]
    nd, nh, nw = (g[0] for g in geo)
    if not _fused3d_ok(L, nw, x.shape[3], x):
        yl, yh = wavedec3(x, w, mode=mode)
        return torch.cat((yl[:, None], torch.stack([yh[k] for k in KEYS3], dim=1)), dim=1)
    B = x.shape[0]
    out = torch.empty((B, 8, nd, nh, nw), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().wdno_dwt3d_analysis(x.data_ptr(), _band_ptrs([out[:, i] for i in range(8)]), 8 * nd * nh * nw, B,
                                              x.shape[1], x.shape[2], x.shape[3], nd, nh, nw, _farr(w.dec_lo[::-1]),
                                              _farr(w.dec_hi[::-1]), L, geo[0][1], _lib.current_stream_ptr()),
               "dwt3d_analysis")
    return out


def dwt2_packed(x, wave, mode):
    """x [B,C,H,W] -> [B,C,4,h,w] = (LL, LH, HL, HH) of one level: `cat((Yl[:, :, None], Yh[0]), 2)` of
    DWTForward(J=1), i.e. Burgers `coef_to_tensor(Yl, Yh)` for J = 1 (wave_trans.py:43-52,107-108) and, for C = 1, the smoke
    builder's `cat((Yl0, Yh0[0][:, 0]), 1)` (wave_trans_2d.py:136-137).  The column pass writes into the packed tensor."""
    _check(x)
    w = _wave(wave)
    L = w.dec_len
    x = x.detach().to(torch.float32).contiguous()
    B, Cc, H, W = x.shape
    nw_, offw, per = _geom(W, L, mode)
    nh_, offh, _ = _geom(H, L, mode)
    tl, th = w.dec_lo[::-1], w.dec_hi[::-1]
    out = torch.empty((B, Cc, 4, nh_, nw_), dtype=torch.float32, device=x.device)
    if _fused2d_ok(L, H, W, nh_, nw_, per, x):
        _ana2d_raw(x, [out[:, :, i] for i in range(4)], tl, th, _geom(H, L, mode), _geom(W, L, mode))
        return out
    lo_w, hi_w = _analysis_raw(x, 3, tl, th, offw, per, nw_)
    _analysis_raw(lo_w, 2, tl, th, offh, per, nh_, lo=out[:, :, 0], hi=out[:, :, 1])
    _analysis_raw(hi_w, 2, tl, th, offh, per, nh_, lo=out[:, :, 2], hi=out[:, :, 3])
    return out
