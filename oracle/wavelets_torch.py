"""TEST INFRASTRUCTURE ONLY -- second, independent CPU restatement of the wavelet libraries WDNO calls, written the
way the libraries themselves work (torch `conv1d` / `conv_transpose1d` with stride 2),
so that it is DIFFERENTIABLE through torch autograd.  Purpose:

  * cross-check of the numpy restatement `oracle/wavelets.py` (loop form) by a convolution form;
  * stand-ins named like the absent third-party packages (`pywt.Wavelet`, `ptwt.wavedec3/waverec3`,
    `pytorch_wavelets.DWTForward/DWTInverse/DWT1DForward/DWT1DInverse`) that `oracle/ref_loader.py` installs into
    `sys.modules`, so the REAL reference glue (`smoke/inference_2d.py`: `guidance_fn`, `InferencePipeline`;
    `burgers/ddpm_burgers/model_utils.py`) runs unchanged in the build container and generates the golden vectors
    of the guided / cascaded sampling paths (tests/golden/make_golden.py).

"parity unpinned" against the real libraries, like oracle/wavelets.py (they are neither vendored nor installable).
"""
import torch
import torch.nn.functional as F
from torch import nn

from .wavelets import filter_bank

KEYS3 = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")


class Wavelet:
    """pywt.Wavelet(name): the four tap lists."""

    def __init__(self, name):
        self.name = name
        dl, dh, rl, rh = filter_bank(name)
        self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi = list(dl), list(dh), list(rl), list(rh)
        self.dec_len = self.rec_len = len(dl)
        self.filter_bank = (self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi)

    def __len__(self):
        return self.dec_len


def dwt_max_level(data_len, filter_len):
    """pywt.dwt_max_level: floor(log2(data_len / (filter_len - 1))), 0 if the filter is longer than the signal; the second
    argument may be a wavelet name (wave_trans.py:87, wave_trans_2d.py:92)"""
    import math
    if isinstance(filter_len, str):
        filter_len = len(filter_bank(filter_len)[0])
    elif hasattr(filter_len, "dec_len"):
        filter_len = filter_len.dec_len
    if filter_len < 2 or data_len < filter_len - 1:
        return 0
    return int(math.floor(math.log2(data_len / (filter_len - 1))))


def _name(w):
    return w if isinstance(w, str) else w.name


def _taps(w, dtype, device):
    dl, dh, rl, rh = filter_bank(_name(w))
    t = lambda a: torch.tensor(a.copy(), dtype=dtype, device=device)
    return t(dl), t(dh), t(rl), t(rh)


def afb1d(x, wave, mode, axis=-1):
    """analysis along `axis`: grouped conv1d, stride 2, weights = reversed dec taps (A.1). -> (lo, hi)"""
    dl, dh, _, _ = _taps(wave, x.dtype, x.device)
    L = dl.numel()
    xm = x.movedim(axis, -1)
    lead = xm.shape[:-1]
    v = xm.reshape(-1, 1, xm.shape[-1])
    N = v.shape[-1]
    w = torch.stack((dl.flip(0), dh.flip(0))).unsqueeze(1)  # [2,1,L]
    if mode == "zero":
        nout = (N + L - 1) // 2
        p = 2 * (nout - 1) - N + L
        if p % 2 == 1:
            v = F.pad(v, (0, 1))
        y = F.conv1d(v, w, stride=2, padding=p // 2)
    elif mode == "periodization":
        if N % 2 == 1:
            v = torch.cat((v, v[..., -1:]), dim=-1)
            N += 1
        v = torch.roll(v, -L // 2, dims=-1)
        y = F.conv1d(v, w, stride=2, padding=L - 1)
        N2 = N // 2
        head = y[..., :L // 2] + y[..., N2:N2 + L // 2]
        y = torch.cat((head, y[..., L // 2:N2]), dim=-1)
    else:
        raise ValueError(mode)
    lo = y[:, 0].reshape(*lead, -1).movedim(-1, axis)
    hi = y[:, 1].reshape(*lead, -1).movedim(-1, axis)
    return lo, hi


def sfb1d(lo, hi, wave, mode, axis=-1):
    """synthesis along `axis`: conv_transpose1d, stride 2, rec taps not reversed (A.2)."""
    _, _, rl, rh = _taps(wave, lo.dtype, lo.device)
    L = rl.numel()
    lm, hm = lo.movedim(axis, -1), hi.movedim(axis, -1)
    lead = lm.shape[:-1]
    n = lm.shape[-1]
    a, b = lm.reshape(-1, 1, n), hm.reshape(-1, 1, n)
    g0, g1 = rl.reshape(1, 1, L), rh.reshape(1, 1, L)
    if mode == "zero":
        y = F.conv_transpose1d(a, g0, stride=2, padding=L - 2) + F.conv_transpose1d(b, g1, stride=2, padding=L - 2)
    elif mode == "periodization":
        y = F.conv_transpose1d(a, g0, stride=2) + F.conv_transpose1d(b, g1, stride=2)
        head = y[..., :L - 2] + y[..., 2 * n:2 * n + L - 2]
        y = torch.cat((head, y[..., L - 2:2 * n]), dim=-1)
        y = torch.roll(y, 1 - L // 2, dims=-1)
    else:
        raise ValueError(mode)
    return y.reshape(*lead, -1).movedim(-1, axis)


class DWTForward(nn.Module):
    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        self.J, self.wave, self.mode = J, _name(wave), mode

    def forward(self, x):
        ll, yh = x, []
        for _ in range(self.J):
            lo_w, hi_w = afb1d(ll, self.wave, self.mode, -1)
            ll, lh = afb1d(lo_w, self.wave, self.mode, -2)
            hl, hh = afb1d(hi_w, self.wave, self.mode, -2)
            yh.append(torch.stack((lh, hl, hh), dim=2))
        return ll, yh


class DWTInverse(nn.Module):
    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        self.wave, self.mode = _name(wave), mode

    def forward(self, coeffs):
        ll, yh = coeffs
        for h in yh[::-1]:
            if ll.shape[-2] > h.shape[-2]:
                ll = ll[..., :-1, :]
            if ll.shape[-1] > h.shape[-1]:
                ll = ll[..., :-1]
            lo = sfb1d(ll, h[:, :, 0], self.wave, self.mode, -2)
            hi = sfb1d(h[:, :, 1], h[:, :, 2], self.wave, self.mode, -2)
            ll = sfb1d(lo, hi, self.wave, self.mode, -1)
        return ll


class DWT1DForward(nn.Module):
    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        self.J, self.wave, self.mode = J, _name(wave), mode

    def forward(self, x):
        lo, his = x, []
        for _ in range(self.J):
            lo, hi = afb1d(lo, self.wave, self.mode, -1)
            his.append(hi)
        return lo, his


class DWT1DInverse(nn.Module):
    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        self.wave, self.mode = _name(wave), mode

    def forward(self, coeffs):
        lo, his = coeffs
        for hi in his[::-1]:
            if lo.shape[-1] > hi.shape[-1]:
                lo = lo[..., :-1]
            lo = sfb1d(lo, hi, self.wave, self.mode, -1)
        return lo


def wavedec3(data, wavelet, *, mode="zero", level=1):
    """ptwt 0.1.6 semantics for level 1 / mode 'zero' (A.3): == three 1-D 'zero' passes along D, H, W."""
    assert level == 1 and mode == "zero"
    out = {}
    for kd, xd in zip("ad", afb1d(data, wavelet, "zero", 1)):
        for kh, xh in zip("ad", afb1d(xd, wavelet, "zero", 2)):
            for kw, xw in zip("ad", afb1d(xh, wavelet, "zero", 3)):
                out[kd + kh + kw] = xw
    return [out["aaa"], {k: out[k] for k in KEYS3}]


def waverec3(coeffs, wavelet):
    b = dict(coeffs[1])
    b["aaa"] = coeffs[0]
    xd = {}
    for kd in "ad":
        xh = {kh: sfb1d(b[kd + kh + "a"], b[kd + kh + "d"], wavelet, "zero", 3) for kh in "ad"}
        xd[kd] = sfb1d(xh["a"], xh["d"], wavelet, "zero", 2)
    return sfb1d(xd["a"], xd["d"], wavelet, "zero", 1)
