"""Offline wavelet-coefficient dataset builders: the bulk users of the DWT kernels.

The reference builds its training sets with the `__main__` blocks of
    smoke/wave_trans_2d.py:61-189   one simulation per iteration: 2 x 3 wavedec3 + 2-D / 1-D transforms, one file per simulation
    burgers/wave_trans.py:66-127    20 000 trajectories per batch: 4 x DWTForward(J=1), one file for the whole set
and reads them back in smoke/ddpm/data_2d.py:156-221 and burgers/ddpm_burgers/data_burgers_1d.py:32-85.  The on-disk
formats (torch.save dictionaries, key names, list lengths, tensor layouts, `shape` as torch.Size) are kept exactly.

B200 form: simulations are batched (`batch_sims` at a time, 13 MB of raw fields each) so one fused 3-D launch transforms
5 x batch_sims fields and writes the sub-bands directly in the packed [.., 8, T', H', W'] layout the files hold
(`wavelets.wavedec3_packed`: no stack / cat pass); one device->host copy per level and kind.  HBM-bound (5.95 MB of
algorithmic traffic per simulation and level-0 transform); in practice the job is bounded by reading
the .npy files.  fp32 CUDA only, like every wdno_b200 path.
"""
import os

import numpy as np
import torch

from . import wavelets as W

SMOKE_KINDS = ("time", "space")


# ------------------------------------------------------------------ smoke (wave_trans_2d.py:61-189)
def smoke_sims_to_coef(X, s, wave_type="bior1.3", mode="zero", N_downsample=3):
    """X [S,5,T,H,W] (rho, v1, v2, c1, c2 of S simulations) and s [S,T] (smoke-out fraction), CUDA fp32 ->
    {kind: {'coef': [S,5,8,T',H',W'] per level, 'init_coef': [S,5,4,H',W'] per level, 'smokeout': [S,2,n] per level}}
    for kind in ('time', 'space'); level i keeps every 2^i-th frame ('time') or row and column ('space')
    (wave_trans_2d.py:124-156).  The smoke-out series follows the frame stride only in the 'time' files (line 153)."""
    S, F, T, H, Wd = X.shape
    X = X.to(torch.float32)
    s = s.to(torch.float32)
    out = {k: dict(coef=[], init_coef=[], smokeout=[]) for k in SMOKE_KINDS}
    with torch.no_grad():
        for i in range(N_downsample):
            st = 2 ** i
            for kind in SMOKE_KINDS:
                X_sub = X[:, :, ::st] if kind == "time" else X[:, :, :, ::st, ::st]
                flat = X_sub.reshape(S * F, *X_sub.shape[2:])  # the strided gather (a view at level 0)
                c = W.wavedec3_packed(flat, wave_type, mode=mode)
                out[kind]["coef"].append(c.view(S, F, *c.shape[1:]))
                first = X_sub[:, :, 0].reshape(S * F, 1, *X_sub.shape[3:])  # frame 0 of every field
                c0 = W.dwt2_packed(first, wave_type, mode)[:, 0]
                out[kind]["init_coef"].append(c0.view(S, F, *c0.shape[1:]))
                s_sub = s[:, None, ::st] if kind == "time" else s[:, None, :]
                lo, hi = W.afb1d(s_sub.contiguous(), wave_type, mode, axis=-1)
                out[kind]["smokeout"].append(torch.cat((lo, hi), dim=1))
    return out


def smoke_max_coef(res, max_coef=None):
    """running per-channel normalisers of wave_trans_2d.py:159-167 (these become RESCALER): int(|coef|.max()) + 1 for the
    40 field/sub-band channels and for the density of the initial frame, int(|smokeout|.max()) for channel 41."""
    m = dict(max_coef) if max_coef is not None else {i: 0 for i in range(5 * 8 + 2)}
    for kind in SMOKE_KINDS:
        for c in res[kind]["coef"]:
            a = c.abs().amax(dim=(0, 3, 4, 5)).cpu()  # [5, 8]
            for j in range(5):
                for i in range(8):
                    m[8 * j + i] = max(int(m[8 * j + i]), int(a[j, i]) + 1)
        for c in res[kind]["init_coef"]:
            m[40] = max(int(m[40]), int(c[:, 0].abs().max()) + 1)
    for c in res["time"]["smokeout"]:
        m[41] = max(int(m[41]), int(c.abs().max()))
    return m


def smoke_load_sim(sim_dir, num_t=32):
    """-> (X [5,num_t,H,W], s [num_t]) float32 CPU from Density/Velocity/Control/Smoke.npy (wave_trans_2d.py:99-109)"""
    def field(name):
        return torch.from_numpy(np.load(os.path.join(sim_dir, name + ".npy"))).float().permute(2, 3, 0, 1)
    X = torch.cat((field("Density"), field("Velocity"), field("Control")), dim=0)[:, :num_t]
    s = torch.from_numpy(np.load(os.path.join(sim_dir, "Smoke.npy"))).float()
    s = (s[:, 1] / s.sum(-1))[:num_t]
    return X, s


def smoke_sim_records(res, j, ori_shape):
    """the two dictionaries the reference saves for simulation j of a batch (wave_trans_2d.py:169-184)"""
    rec = {}
    for kind in SMOKE_KINDS:
        r = res[kind]
        rec[kind] = {"coef": [c[j].clone() for c in r["coef"]],
                     "init_coef": [c[j].clone() for c in r["init_coef"]],
                     "smokeout": [c[j].clone() for c in r["smokeout"]],
                     "shape": [c.shape[-3:] for c in r["coef"]],
                     "ori_shape": torch.Size(ori_shape)}
    return rec


def build_smoke_coef_files(root="./data/2d/", dirname="train/", sim_range=range(20000), wave_type="bior1.3", mode="zero",
                           N_downsample=3, num_t=32, batch_sims=32, device="cuda"):
    """the job of wave_trans_2d.py's __main__: <root>/<dirname>/sim_%06d/*.npy ->
    <root>/<dirname>/<wave>_<mode>/{time,space}_downsample/%06d.  Returns the max_coef list the reference prints."""
    wave_dir = os.path.join(root, dirname, "{}_{}/".format(wave_type, mode))
    for kind in SMOKE_KINDS:
        os.makedirs(os.path.join(wave_dir, kind + "_downsample"), exist_ok=True)
    max_coef = None
    ids = list(sim_range)
    for b0 in range(0, len(ids), batch_sims):
        chunk = ids[b0:b0 + batch_sims]
        sims = [smoke_load_sim(os.path.join(root, dirname, "sim_{:06d}".format(i)), num_t) for i in chunk]
        X = torch.stack([x for x, _ in sims]).pin_memory().to(device, non_blocking=True)
        s = torch.stack([t for _, t in sims]).pin_memory().to(device, non_blocking=True)
        res = smoke_sims_to_coef(X, s, wave_type, mode, N_downsample)
        max_coef = smoke_max_coef(res, max_coef)
        host = {k: {n: [c.cpu() for c in v] for n, v in r.items()} for k, r in res.items()}
        for j, sim_id in enumerate(chunk):
            rec = smoke_sim_records(host, j, X.shape[2:])
            for kind in SMOKE_KINDS:
                torch.save(rec[kind], os.path.join(wave_dir, kind + "_downsample/", "{:06d}".format(sim_id)))
    return list(max_coef.values()) if max_coef is not None else None


# ------------------------------------------------------------------ burgers (wave_trans.py:66-127)
def burgers_data_to_coef(data, wave_type="bior2.4", mode="periodization", N_downsample=4):
    """data [N,2,nt,nx] (u, f with the zero last row appended; CUDA fp32) ->
    {'coef': [[N,2,4,h_i,w_i] per level], 'shape': [torch.Size([4,h_i,w_i])] (the reference stores shape[2:] of the
    packed tensor, band axis included), 'ori_shape': torch.Size([nt,nx])}
    (wave_trans.py:103-120): level i keeps every 2^i-th row and column, then one DWT level, packed (LL,LH,HL,HH)."""
    data = data.to(torch.float32)
    coef = []
    with torch.no_grad():
        for i in range(N_downsample):
            st = 2 ** i
            coef.append(W.dwt2_packed(data[:, :, ::st, ::st], wave_type, mode))
    return {"coef": coef, "shape": [c.shape[2:] for c in coef], "ori_shape": data.shape[2:]}


def build_burgers_coef_file(train_path="data/1d/train", out_path=None, wave_type="bior2.4", mode="periodization",
                            N_downsample=4, batch_size=20000, device="cuda"):
    """the job of wave_trans.py's __main__: {'u': [N,nt+1,nx], 'f': [N,nt,nx]} -> data/1d/coef_<wave>_<mode>_super"""
    all_data = torch.load(train_path)
    u, f = all_data["u"], all_data["f"]
    f = torch.cat((f, torch.zeros(u.shape[0], 1, f.shape[-1])), dim=1)
    data = torch.cat((u.unsqueeze(1), f.unsqueeze(1)), dim=1).float()
    parts = []
    for b0 in range(0, data.shape[0], batch_size):
        r = burgers_data_to_coef(data[b0:b0 + batch_size].pin_memory().to(device, non_blocking=True), wave_type, mode,
                                 N_downsample)
        parts.append([c.cpu() for c in r["coef"]])
    coef = [torch.cat([p[i] for p in parts]) for i in range(N_downsample)]
    out = {"coef": coef, "shape": [c.shape[2:] for c in coef], "ori_shape": data.shape[2:]}
    if out_path is None:
        out_path = os.path.join(os.path.dirname(train_path), "coef_{}_{}_super".format(wave_type, mode))
    torch.save(out, out_path)
    return out_path
