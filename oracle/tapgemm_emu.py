"""TEST INFRASTRUCTURE ONLY -- numpy emulation of csrc/tapgemm.cu's addressing semantics.

Walks the same unit / K-set / plane / tap / N-chunk tables the CUDA kernel walks (built by
wdno_b200.tapgemm.TapGemm on the CPU) so that weight packing, tap shifts, space-to-depth /
depth-to-space phases and padded-row geometry can be checked against torch.nn.functional
convolutions without a GPU.  Not imported by the product path.
"""
import ctypes as C

import numpy as np
import torch

from wdno_b200._lib import KSet, NChunk, Tap


def _table(t, typ):
    raw = bytes(t.cpu().numpy().tobytes())
    n = len(raw) // C.sizeof(typ)
    return (typ * n).from_buffer_copy(raw)


def emulate(plan, src0, src1=None, coef0=None, coef1=None, resid=None, groups=8, want_stats=False,
            out_fp32_bfchw=False):
    """plan: TapGemm built with device='cpu'.  src*: fp16 [B,D,Hs,Ws,C] CPU tensors.  Returns (out, stats)."""
    B, D, Hs, Ws, _ = src0.shape
    if plan.kind in ("down144", "unshuffle"):
        H, W = Hs // 2, Ws // 2
    elif plan.kind == "conv" and plan.up2:
        H, W = Hs * 2, Ws * 2
    else:
        H, W = Hs, Ws
    p = plan._plan(B, D, H, W)
    Wfull, W = W, p.W            # p.W: tap-grid columns of one strip (== Wfull when p.strips == 1)
    B_log, D_log = B, D
    B, D = p.B, p.D              # batch folded into depth planes for 2-D layers (p.fold): B*D unchanged
    assert B * D == B_log * D_log and (p.fold or (B, D) == (B_log, D_log))
    pk = plan._packed[(p.KC, bool(p.zstack))]
    chunks = _table(pk["chunks"], NChunk)
    sets = _table(pk["sets"], KSet)
    taps = _table(pk["taps_dev"][p.Wp], Tap)
    wpk = pk["wpacked"].float().numpy()
    KC, N = p.KC, p.N
    rows = N * (p.KD if p.zstack else 1)   # rows of one weight tile (kz-stacked tiles hold [kz][N])
    tile_elems = rows * KC
    fold5 = lambda t: None if t is None else t.float().numpy().reshape((B, D) + tuple(t.shape[2:]))
    srcs = [fold5(src0), fold5(src1)]
    coefs = [coef0, coef1]
    NACC = p.ZT * p.PT
    P = p.ZT + p.KD - 1
    S = 128 * p.PT + p.maxshift
    positions = H * p.Wp
    ptiles = (positions + 128 * p.PT - 1) // (128 * p.PT)
    zgroups = (D + p.ZT - 1) // p.ZT
    Ho, Wo = (2 * H, 2 * Wfull) if plan.kind == "up144" else (H, Wfull)
    cout = plan.cout
    if out_fp32_bfchw:
        out = np.zeros((B, D, cout, H, Wfull), np.float32)
    else:
        out = np.zeros((B, D, Ho, Wo, cout), np.float32)
    stats = np.zeros((B_log, groups, 2), np.float64)
    bias = None if plan.bias is None else plan.bias.cpu().numpy()
    res = fold5(resid)
    cpg = max(1, cout // groups)
    for bs in range(B * p.strips):
        b, xs0 = bs // p.strips, (bs % p.strips) * W
        for zg in range(zgroups):
            for pt in range(ptiles):
                for nc in range(p.n_chunks):
                    ci = chunks[nc]
                    o0, z0 = pt * 128 * p.PT, zg * p.ZT
                    acc = np.zeros((NACC, 128, N), np.float64)
                    tile_i = ci.w_tile_off
                    for si in range(ci.set_count):
                        st = sets[ci.set_begin + si]
                        src = srcs[st.src]
                        # build the P slabs [S, KC]
                        slabs = np.zeros((P, S, KC), np.float32)
                        for j in range(P):
                            zi = z0 - p.pz + j
                            if zi < 0 or zi >= D:
                                continue
                            q = o0 + np.arange(S)
                            yp, xp = q // p.Wp, q % p.Wp
                            y, x = yp - p.py, xp - p.px + xs0
                            ok = (y >= 0) & (y < H) & (x >= 0) & (x < Wfull)
                            ys, xs = y.copy(), x.copy()
                            if p.src_mode == 1:
                                ys, xs = 2 * y + st.ph_y, 2 * x + st.ph_x
                            elif p.src_mode == 2:
                                ys, xs = y >> 1, x >> 1
                            ysc, xsc = np.where(ok, ys, 0), np.where(ok, xs, 0)
                            v = src[b, zi, ysc, xsc, st.ch_off:st.ch_off + KC]
                            if coefs[st.src] is not None:
                                a, c = coefs[st.src]
                                smp = b * D + zi if p.fold else b
                                a = a[smp, st.ch_off:st.ch_off + KC].numpy()
                                c = c[smp, st.ch_off:st.ch_off + KC].numpy()
                                t = a * v + c
                                v = t / (1.0 + np.exp(-t))
                                v = v.astype(np.float16).astype(np.float32)
                            slabs[j] = np.where(ok[:, None], v, 0.0)
                        for t in range(st.tap_count):
                            tp = taps[st.tap_begin + t]
                            tile = wpk[tile_i * tile_elems:(tile_i + 1) * tile_elems].reshape(KC // 8, rows, 8)
                            wmat = tile.transpose(1, 0, 2).reshape(rows, KC)  # [rows, KC]
                            tile_i += 1
                            if p.zstack:
                                # mma_role_zstack: input plane j feeds output plane o = j - kz through rows [kz*N, (kz+1)*N)
                                for j in range(P):
                                    a_rows = slabs[j, tp.shift: tp.shift + 128].astype(np.float64)
                                    for kz in range(max(0, j - (p.ZT - 1)), min(p.KD - 1, j) + 1):
                                        acc[j - kz] += a_rows @ wmat[kz * N:(kz + 1) * N].T.astype(np.float64)
                                continue
                            for za in range(p.ZT):
                                for pi in range(p.PT):
                                    a_rows = slabs[tp.kz + za, tp.shift + pi * 128: tp.shift + pi * 128 + 128]
                                    acc[za * p.PT + pi] += a_rows.astype(np.float64) @ wmat.T.astype(np.float64)
                    assert tile_i - ci.w_tile_off == ci.n_tiles
                    # epilogue
                    for a in range(NACC):
                        za, pi = a // p.PT, a % p.PT
                        z = z0 + za
                        if z >= D:
                            continue
                        for r in range(128):
                            o = o0 + pi * 128 + r
                            y, xl = o // p.Wp, o % p.Wp
                            x = xl + xs0
                            if y >= H or xl >= W or x >= Wfull:
                                continue
                            v = acc[a, r, :ci.n_valid].astype(np.float32)
                            ch = slice(ci.out_ch_off, ci.out_ch_off + ci.n_valid)
                            if bias is not None:
                                v = v + bias[ch]
                            if want_stats:
                                for k in range(ci.n_valid):
                                    g = (ci.out_ch_off + k) // cpg
                                    smp = b * D + z if p.fold else b
                                    stats[smp, g, 0] += v[k]
                                    stats[smp, g, 1] += float(v[k]) ** 2
                            if out_fp32_bfchw:
                                out[b, z, ch, y, x] = v
                            elif plan.kind == "up144":
                                yy, xx = 2 * y + ci.ph_y, 2 * x + ci.ph_x
                                if res is not None:
                                    v = v + res[b, z, yy, xx, ch]
                                out[b, z, yy, xx, ch] = v
                            else:
                                if res is not None:
                                    v = v + res[b, z, y, x, ch]
                                out[b, z, y, x, ch] = v
    out = out.reshape((B_log, D_log) + out.shape[2:])
    return torch.from_numpy(out), torch.from_numpy(stats)
