"""Coefficient <-> tensor packing helpers (pure index permutations; bit-exact by construction).

smoke   : /root/reference/smoke/wave_trans_2d.py:17-58, smoke/ddpm/wave_utils.py:1-14
burgers : /root/reference/burgers/wave_trans.py:18-62, burgers/ddpm_burgers/wavelet_utils.py:5-28
"""
import torch
import torch.nn.functional as F

SMOKE_KEYS = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")
N_FIELDS = 5  # rho, v1, v2, c1, c2


# ------------------------------------------------------------------ smoke (3-D transform, 5 fields x 8 sub-bands)
def smoke_tensor_to_coef(coef_tensor, shape, upsample_type=None):
    """[N, >=40, T', H', W'] (channel 8*field + band) -> (Yl [5N, T, H, W], {7 detail bands [5N, T, H, W]}).
    'time' / 'space' variants were replicate-padded by one coefficient, so their crop starts at index 1."""
    T, H, W = shape[-3], shape[-2], shape[-1]
    t0 = 1 if upsample_type == "time" else 0
    s0 = 1 if upsample_type == "space" else 0
    if upsample_type not in (None, "time", "space"):
        raise ValueError(upsample_type)
    Hc = shape[-2]  # the reference crops rows and columns with the same bounds (shape[-2])
    crop = coef_tensor[:, :8 * N_FIELDS, t0:t0 + T, s0:s0 + Hc, s0:s0 + Hc]
    n = crop.shape[0]
    per_field = crop.reshape(n, N_FIELDS, 8, *crop.shape[2:])
    yl = per_field[:, :, 0].reshape(-1, T, H, W)
    det = per_field[:, :, 1:].reshape(-1, 7, T, H, W)
    return yl, {k: det[:, i] for i, k in enumerate(SMOKE_KEYS)}


def smoke_coef_to_tensor(coef, pad=False):
    yl, yh = coef[0], coef[1]
    return torch.cat((yl[:, None], torch.stack(list(yh.values()), dim=1)), dim=1)


def smoke_upsample_coef(w_sub, shape, type):
    """nearest x2 along time ('time') or along both space axes (anything else); w_sub [N, nt, layer, nx, nx]"""
    if type == "time":
        return w_sub.repeat_interleave(2, dim=1)
    return w_sub.repeat_interleave(2, dim=3).repeat_interleave(2, dim=4)


# ------------------------------------------------------------------ burgers (2-D transform, u and f)
def burgers_tensor_to_coef(coef_tensor, shape):
    """[N, >=8, Hp, Wp] (u: LL,LH,HL,HH ; f: LL,LH,HL,HH) -> (Yl [N,2,H,W], [Yh [N,2,3,H,W]])"""
    H, W = shape[-2], shape[-1]
    c = coef_tensor[:, :8, :H, :W].reshape(coef_tensor.shape[0], 2, 4, H, W)
    return c[:, :, 0], [c[:, :, 1:4]]


burgers_tensor_to_coef_super = burgers_tensor_to_coef


def burgers_coef_to_tensor(Yl, Yh, pad=False):
    """(Yl [N,C,h,w], [Yh_i [N,C,3,h_i,w_i]] finest first) -> [N, C, 1+3J, H, W]: level i replicated 2^i times
    (nearest), rows padded by repeating the last row up to the finest grid's H + 2^(J-1) - 1"""
    J = len(Yh)
    H = Yh[0].shape[-2] + 2 ** (J - 1) - 1
    W = Yh[0].shape[-1]
    out = torch.zeros(Yl.shape[0], Yl.shape[1], 1 + 3 * J, H, W, device=Yl.device, dtype=Yl.dtype)
    r = 2 ** (J - 1)
    out[:, :, 0] = Yl.repeat_interleave(r, dim=-2).repeat_interleave(r, dim=-1)
    for i in range(J):
        r = 2 ** i
        rep = Yh[i].repeat_interleave(r, dim=-2).repeat_interleave(r, dim=-1)
        extra = 2 ** (J - 1) - 2 ** i
        if extra > 0:
            rep = torch.cat((rep, rep[:, :, :, -1:].expand(-1, -1, -1, extra, -1)), dim=3)
        out[:, :, 1 + 3 * i:4 + 3 * i] = rep
    if pad:
        ut, ux = int(out.shape[-2] / 40), int(out.shape[-1] / 60)
        out = F.pad(out, (0, 64 * ux - out.shape[-1], 0, 64 * ut - out.shape[-2]), "constant", 0)
    return out


def burgers_upsample_coef(w_sub, shape):
    """nearest x2 in both axes; w_sub [N, layer, nt, nx]"""
    return w_sub.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def burgers_get_wt_T(test_data, shape):
    out = [test_data[:, 0, shape[-1][-2] - 1]]
    for i in range(len(shape)):
        out.append(test_data[:, 1 + 3 * i:4 + 3 * i, shape[i][-2] - 1])
    return out
