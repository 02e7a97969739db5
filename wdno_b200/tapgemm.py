"""Host-side plans for the tcgen05 tap-GEMM kernel (csrc/tapgemm.cu, C ABI: wdno_tapgemm).

A `TapGemm` owns the fp16 weight tiles of one dense layer, re-packed from the reference's
`state_dict` tensor into the kernel's tile order, plus the tap / K-set / N-chunk tables.  Calling it
enqueues one kernel on the current torch CUDA stream.  Layer kinds (reference call sites):

  conv       nn.Conv3d / nn.Conv2d / nn.Linear with "same" padding or 1x1
             (video_diffusion_pytorch_conv3d.py:192,216,238-239,291-292,393,471; unet.py:133,162,190-192,233-234,317,369)
  down144    nn.Conv3d(dim, dim, (1,4,4), (1,2,2), (0,1,1))            (conv3d.py:162-163)
  up144      nn.ConvTranspose3d(dim, dim, (1,4,4), (1,2,2), (0,1,1))   (conv3d.py:159-160)
  unshuffle  Rearrange('b c (h p1) (w p2) -> b (c p1 p2) h w') + Conv2d(4c, c', 1)   (unet.py:41-45)
  conv+up2   nn.Upsample(scale_factor=2, 'nearest') + Conv2d(3, padding=1)           (unet.py:35-39): kind='conv', up2=True
"""
import ctypes as C
import os

import torch

from . import _lib, _timing
from ._lib import KSet, NChunk, Tap, TapGemmParams

_SMEM_LIMIT = 227 * 1024
_BAR_BYTES = 512 + 384 * 8 + 4 * 32 * 144 + 1024 * 4 + 32 * 40 + 128 * 24  # barriers, tap table, epilogue staging, bias, N-chunk / K-set tables


def _device_bytes(ctypes_array, device):
    raw = bytes(ctypes_array)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    return t.to(device)


_CLUSTER_CAP = {}


def _cluster_cta_cap(smem_bytes):
    """CTAs of 2-CTA clusters resident at once (cached per shared-memory size; 0 when the query fails)"""
    key = int(smem_bytes)
    if key not in _CLUSTER_CAP:
        n = _lib.lib().wdno_tapgemm_max_cluster_ctas(key)
        _CLUSTER_CAP[key] = max(0, int(n))
    return _CLUSTER_CAP[key]


def _round_up(a, b):
    return (a + b - 1) // b * b


class TapGemm:
    def __init__(self, weight, bias=None, *, kind="conv", src_channels=None, up2=False,
                 n_tile=None, device=None, kc=None):
        """weight: fp32 tensor in the reference layout of the layer kind (see module docstring)."""
        # packing runs on the device the weight lives on (training refreshes the tiles from live CUDA parameters every step;
        # a CPU weight is packed on the host and uploaded)
        w = weight.detach().to(torch.float32)
        self.kind = kind
        self.up2 = bool(up2)
        self.device = torch.device(device if device is not None else "cuda")
        if w.dim() == 2:  # nn.Linear [out, in]
            w = w[:, :, None, None, None]
        elif w.dim() == 4:  # nn.Conv2d [out, in, kh, kw]
            w = w[:, :, None]
        assert w.dim() == 5
        if kind == "up144":
            self.cin, self.cout = w.shape[0], w.shape[1]
        elif kind == "unshuffle":
            self.cout, self.cin = w.shape[0], w.shape[1] // 4
        else:
            self.cout, self.cin = w.shape[0], w.shape[1]
        self.src_channels = tuple(src_channels) if src_channels is not None else (self.cin,)
        assert sum(self.src_channels) >= self.cin
        self.w = w
        self.bias = None
        if kind == "conv":
            self.KD, self.KH, self.KW = w.shape[2:]
        elif kind in ("down144", "up144"):
            assert tuple(w.shape[2:]) == (1, 4, 4)
            self.KD, self.KH, self.KW = 1, 3, 3
        elif kind == "unshuffle":
            assert tuple(w.shape[2:]) == (1, 1, 1)
            self.KD, self.KH, self.KW = 1, 1, 1
        else:
            raise ValueError(kind)
        assert self.KD % 2 == 1 and self.KH % 2 == 1 and self.KW % 2 == 1
        if n_tile is None:
            n_tile = 64 if self.cout >= 64 else _round_up(self.cout, 16)
            big = int(os.environ.get("WDNO_NTILE_BIG", "1"))  # tuning knob: N=128 tiles for wide spatial convolutions
            if big and self.cout % 128 == 0 and kind == "conv" and w.shape[2] * w.shape[3] * w.shape[4] > 1:
                n_tile = 128
        self.N = int(n_tile)
        self.cout_pad = _round_up(self.cout, self.N)
        self.kc_override = kc
        if bias is not None:
            b = torch.zeros(self.cout_pad, dtype=torch.float32, device=self.device)
            b[: self.cout] = bias.detach().float().to(self.device)
            self.bias = b
        self._packed = {}   # KC -> dict(wpacked, chunks, sets, taps, n_chunks, tables kept alive)
        self._launch = {}   # geometry key -> (params, keepalive)
        # plain 1x1 layers (no fused GroupNorm prologue / statistics) run on the HBM-bound mma.sync kernel csrc/conv1x1.cu
        self._c1 = None
        if (kind == "conv" and (self.KD, self.KH, self.KW) == (1, 1, 1) and len(self.src_channels) <= 2
                and all(c % 32 == 0 for c in self.src_channels) and os.environ.get("WDNO_CONV1X1", "1") != "0"
                and self.device.type == "cuda"):
            npad = _round_up(self.cout, 64)
            b1 = None
            if bias is not None:
                b1 = torch.zeros(npad, dtype=torch.float32, device=self.device)
                b1[: self.cout] = bias.detach().float().to(self.device)
            self._c1 = dict(w=self._c1_weight(npad), bias=b1, npad=npad)

    def _c1_weight(self, npad):
        w1 = torch.zeros(npad, sum(self.src_channels), dtype=torch.float32, device=self.w.device)
        w1[: self.cout, : self.cin] = self.w[:, :, 0, 0, 0]
        return w1.to(torch.float16).to(self.device)

    def refresh(self, weight, bias=None):
        """Re-pack the tiles of every plan built so far from a new weight tensor of the same shape, IN PLACE (device
        addresses baked into captured graphs and launch structs stay valid).  Used by the training step after each
        optimiser update and by engines whose parameters changed in place."""
        w = weight.detach().to(torch.float32)
        if w.dim() == 2:
            w = w[:, :, None, None, None]
        elif w.dim() == 4:
            w = w[:, :, None]
        assert w.shape == self.w.shape
        self.w = w
        for (KC, zstack), pk in self._packed.items():
            pk["wpacked"].copy_(self._pack_weights(KC, zstack)[0].to(torch.float16))
        if self.bias is not None and bias is not None:
            self.bias[: self.cout].copy_(bias.detach().float())
        if self._c1 is not None:
            self._c1["w"].copy_(self._c1_weight(self._c1["npad"]))
            if self._c1["bias"] is not None and bias is not None:
                self._c1["bias"][: self.cout].copy_(bias.detach().float())

    # ------------------------------------------------------------------ packing
    def _virtual_cin(self):
        """total channels of the (concatenated) sources"""
        return sum(self.src_channels)

    def _pack(self, KC, zstack=False):
        if (KC, zstack) in self._packed:
            return self._packed[(KC, zstack)]
        wpacked, taps, sets, chunks = self._pack_weights(KC, zstack)
        sets_c = (KSet * len(sets))(*[KSet(s["src"], s["ch_off"], s["ph_y"], s["ph_x"], s["tap_begin"], s["tap_count"])
                                       for s in sets])
        chunks_c = (NChunk * len(chunks))(*[NChunk(c["out_ch_off"], c["n_valid"], c["ph_y"], c["ph_x"], c["set_begin"],
                                                   c["set_count"], c["n_tiles"], 0, c["w_tile_off"]) for c in chunks])
        pk = dict(
            wpacked=wpacked.to(torch.float16).to(self.device),
            sets=_device_bytes(sets_c, self.device),
            chunks=_device_bytes(chunks_c, self.device),
            taps_kyx=taps, n_chunks=len(chunks), n_sets=len(sets),
            taps_dev={},  # Wp -> device tap table
        )
        self._packed[(KC, zstack)] = pk
        return pk

    def _pack_weights(self, KC, zstack=False):
        """-> (fp32 weight tiles in kernel order, tap list, K-set list, N-chunk list): a pure function of self.w"""
        N, kind = self.N, self.kind
        w = self.w
        ctot = self._virtual_cin()
        assert ctot % KC == 0, f"source channels {ctot} not a multiple of KC={KC}"
        for c in self.src_channels[:-1]:
            assert c % KC == 0, "concat boundary must be KC-aligned"
        nck = ctot // KC            # channel chunks
        ncn = self.cout_pad // N    # cout chunks

        def src_of(chunk):
            ch = chunk * KC
            for si, c in enumerate(self.src_channels):
                if ch < c:
                    return si, ch
                ch -= c
            raise AssertionError

        taps, sets, chunks, tiles = [], [], [], []

        def tile_from(wslice):
            """wslice: [N, KC] fp32 (rows = output channels) -> [KC/8, N, 8]"""
            return wslice.reshape(N, KC // 8, 8).permute(1, 0, 2).contiguous()

        def padded_w(wfull):
            """[cout, cin, ...] -> zero-padded to [cout_pad, ctot, ...]"""
            out = torch.zeros((self.cout_pad, ctot) + tuple(wfull.shape[2:]), dtype=torch.float32, device=wfull.device)
            out[: wfull.shape[0], : wfull.shape[1]] = wfull
            return out

        if kind == "conv" and zstack:
            # kz-stacked tiles: one tile per (N-chunk, K-set, in-plane tap) with rows [kz][N] (csrc/tapgemm.cu, mma_role_zstack)
            wp = padded_w(w)
            KD, KH, KW = self.KD, self.KH, self.KW
            T = KH * KW
            for ky in range(KH):
                for kx in range(KW):
                    taps.append((0, ky, kx))
            for ck in range(nck):
                s_, ch = src_of(ck)
                sets.append(dict(src=s_, ch_off=ch, ph_y=0, ph_x=0, tap_begin=0, tap_count=T))
            # [NC, N, nck, KC/8, 8, KD, T] -> [NC, nck, T, KC/8, KD, N, 8]
            wr = wp.reshape(ncn, N, nck, KC // 8, 8, KD, T).permute(0, 2, 6, 3, 5, 1, 4).contiguous()
            wpacked = wr.reshape(-1)
            for cn in range(ncn):
                chunks.append(dict(out_ch_off=cn * N, n_valid=min(N, self.cout - cn * N), ph_y=0, ph_x=0,
                                   set_begin=0, set_count=nck, n_tiles=nck * T, w_tile_off=cn * nck * T))
        elif kind == "conv":
            wp = padded_w(w)
            KD, KH, KW = self.KD, self.KH, self.KW
            tap_list = [(kz, ky, kx) for kz in range(KD) for ky in range(KH) for kx in range(KW)]
            # one shared tap table (shift filled at launch: depends on Wp) -> store (kz,ky,kx)
            for (kz, ky, kx) in tap_list:
                taps.append((kz, ky, kx))
            for ck in range(nck):
                s, ch = src_of(ck)
                sets.append(dict(src=s, ch_off=ch, ph_y=0, ph_x=0, tap_begin=0, tap_count=len(tap_list)))
            # vectorised packing: [NC, N, nck, KC/8, 8, T] -> [NC, nck, T, KC/8, N, 8]
            T = len(tap_list)
            wr = wp.reshape(ncn, N, nck, KC // 8, 8, T).permute(0, 2, 5, 3, 1, 4).contiguous()
            wpacked = wr.reshape(-1)
            per_chunk_tiles = nck * T
            for cn in range(ncn):
                chunks.append(dict(out_ch_off=cn * N, n_valid=min(N, self.cout - cn * N), ph_y=0, ph_x=0,
                                   set_begin=0, set_count=nck, n_tiles=per_chunk_tiles,
                                   w_tile_off=cn * per_chunk_tiles))
        elif kind == "down144":
            wp = padded_w(w)  # [cout_pad, ctot, 1, 4, 4]
            # phase a (row parity of the source pixel), taps (tky -> ky):  a=0: {1:1, 2:3} ; a=1: {0:0, 1:2}
            ph = {0: [(1, 1), (2, 3)], 1: [(0, 0), (1, 2)]}
            tile_list = []
            for a in (0, 1):
                for b in (0, 1):
                    tb = len(taps)
                    pairs = [(tky, ky, tkx, kx) for (tky, ky) in ph[a] for (tkx, kx) in ph[b]]
                    for (tky, ky, tkx, kx) in pairs:
                        taps.append((0, tky, tkx))
                    for ck in range(nck):
                        s, ch = src_of(ck)
                        sets.append(dict(src=s, ch_off=ch, ph_y=a, ph_x=b, tap_begin=tb, tap_count=4,
                                         _pairs=pairs, _ck=ck))
            per_chunk = 0
            for cn in range(ncn):
                for st in sets:
                    for (tky, ky, tkx, kx) in st["_pairs"]:
                        ck = st["_ck"]
                        tile_list.append(tile_from(wp[cn * N:(cn + 1) * N, ck * KC:(ck + 1) * KC, 0, ky, kx]))
                if cn == 0:
                    per_chunk = len(tile_list)
                chunks.append(dict(out_ch_off=cn * N, n_valid=min(N, self.cout - cn * N), ph_y=0, ph_x=0,
                                   set_begin=0, set_count=len(sets), n_tiles=per_chunk, w_tile_off=cn * per_chunk))
            wpacked = torch.stack(tile_list).reshape(-1)
        elif kind == "up144":
            # w: [cin, cout, 1, 4, 4]; out phase A: taps (tky -> ky): A=0: {1:1, 0:3} ; A=1: {2:0, 1:2}
            wt = torch.zeros((self.cout_pad, ctot, 4, 4), dtype=torch.float32, device=w.device)
            wt[: self.cout, : self.cin] = w[:, :, 0].permute(1, 0, 2, 3)
            ph = {0: [(0, 3), (1, 1)], 1: [(1, 2), (2, 0)]}
            tile_list = []
            for A in (0, 1):
                for Bp in (0, 1):
                    pairs = [(tky, ky, tkx, kx) for (tky, ky) in ph[A] for (tkx, kx) in ph[Bp]]
                    tb = len(taps)
                    for (tky, ky, tkx, kx) in pairs:
                        taps.append((0, tky, tkx))
                    sb = len(sets)
                    for ck in range(nck):
                        s, ch = src_of(ck)
                        sets.append(dict(src=s, ch_off=ch, ph_y=0, ph_x=0, tap_begin=tb, tap_count=4))
                    for cn in range(ncn):
                        t0 = len(tile_list)
                        for ck in range(nck):
                            for (tky, ky, tkx, kx) in pairs:
                                tile_list.append(tile_from(wt[cn * N:(cn + 1) * N, ck * KC:(ck + 1) * KC, ky, kx]))
                        chunks.append(dict(out_ch_off=cn * N, n_valid=min(N, self.cout - cn * N), ph_y=A, ph_x=Bp,
                                           set_begin=sb, set_count=nck, n_tiles=len(tile_list) - t0, w_tile_off=t0))
            wpacked = torch.stack(tile_list).reshape(-1)
        elif kind == "unshuffle":
            # w: [cout, 4*cin, 1,1,1] with input channel index c*4 + p1*2 + p2
            w4 = w[:, :, 0, 0, 0].reshape(self.cout, self.cin, 2, 2)
            wp = torch.zeros((self.cout_pad, ctot, 2, 2), dtype=torch.float32, device=w.device)
            wp[: self.cout, : self.cin] = w4
            taps.append((0, 0, 0))
            tile_list = []
            for a in (0, 1):
                for b in (0, 1):
                    for ck in range(nck):
                        s, ch = src_of(ck)
                        sets.append(dict(src=s, ch_off=ch, ph_y=a, ph_x=b, tap_begin=0, tap_count=1, _ck=ck))
            per_chunk = len(sets)
            for cn in range(ncn):
                for st in sets:
                    ck = st["_ck"]
                    tile_list.append(tile_from(wp[cn * N:(cn + 1) * N, ck * KC:(ck + 1) * KC, st["ph_y"], st["ph_x"]]))
                chunks.append(dict(out_ch_off=cn * N, n_valid=min(N, self.cout - cn * N), ph_y=0, ph_x=0,
                                   set_begin=0, set_count=len(sets), n_tiles=per_chunk, w_tile_off=cn * per_chunk))
            wpacked = torch.stack(tile_list).reshape(-1)
        else:
            raise ValueError(kind)

        return wpacked, taps, sets, chunks

    # ------------------------------------------------------------------ geometry / smem plan
    def _plan(self, B, D, H, W):
        """(B, D, H, W) = tap grid (== output grid before depth-to-space).  Chooses the accumulator shape (ZT x PT),
        K-set width KC, slab ring depth, weight stages (TPS tiles per bulk copy) and whether 1x1 layers keep their
        slabs resident across N-chunks, all under the 227 KB shared-memory budget."""
        key = (B, D, H, W)
        if key in self._launch:
            return self._launch[key]
        # Column strips: a wide plane with a large in-plane kernel (7x7 on 80x80) needs a haloed slab of
        # 128 + (KH-1)*Wp + KW-1 positions per 128 outputs -- too big for the kz-stacked plan's ZT+KD-1 resident planes.
        # Cutting the plane into strips of ~40 columns (each a work-item dimension with real neighbour columns as halo)
        # shrinks Wp and brings the stacked plan back (measured on the 82->64 7^3 stem at 80x80: 261 -> see DESIGN.md).
        cands = [1]
        if (self.kind == "conv" and not self.up2 and self.KD in (3, 7) and self.N == 64 and D >= 4 and self.KH * self.KW > 9
                and int(os.environ.get("WDNO_STRIPS", "1"))):
            cands += [s_ for s_ in (2, 3, 4) if (W + s_ - 1) // s_ >= 32]
        if os.environ.get("WDNO_FORCE_STRIPS"):
            cands = [int(os.environ["WDNO_FORCE_STRIPS"])]
        chosen = None
        for strips in cands:
            p = self._plan_strips(B, D, H, W, strips)
            if chosen is None:
                chosen = p
            if p is not None and (p.zstack or os.environ.get("WDNO_FORCE_STRIPS")):
                chosen = p
                break
        # Batch folding (2-D layers, D == 1): 4 (or 2) samples become the depth planes of one "sample" so that ZT of them
        # share every weight tile -- at 8x8 / 16x16 resolution with 512..1024 channels a work item otherwise streams
        # megabytes of weights from L2 for one 128-position tile (measured 194 TFLOP/s on 1024->1024 3x3 at 8x8).
        if (D == 1 and self.KD == 1 and chosen is not None and not chosen.zstack
                and os.environ.get("WDNO_FOLD", "1") != "0"):
            est0 = self._last_est
            force = os.environ.get("WDNO_FOLD") == "force"
            for f in (4, 2):
                if B % f:
                    continue
                pf = self._plan_strips(B // f, f, H, W, 1, fold=True)
                if pf is None:
                    break
                if chosen.reuse:
                    # A-stationary 1x1 plans carry no estimate: more planes per item = fewer weight re-streams
                    take = bool(pf.reuse) and pf.ZT > chosen.ZT
                else:
                    # clear wins only, and only planes that fit one 128-position tile (8x8): measured on the Burgers U-Net,
                    # folding 16x16 / 32x32 layers (ZT = 4, PT = 1) loses 10-30 % against ZT = 1, PT = 4
                    take = (est0 is not None and self._last_est is not None and self._last_est < 0.8 * est0
                            and H * (W + self.KW // 2) <= 128)
                if take or force:
                    chosen = pf
                break
        if chosen is None:
            raise ValueError(f"tapgemm: no shared-memory plan for grid {(B, D, H, W)} taps {(self.KD, self.KH, self.KW)}")
        self._launch[key] = chosen
        return chosen

    def _plan_strips(self, B, D, H, Wfull, strips, fold=False):
        KD, KH, KW = self.KD, self.KH, self.KW
        self._last_est = None
        two_d = fold or D == 1
        if strips == 1:
            # padded row width: ONE shared run of KW//2 zero columns per row -- the left pad of row y+1 doubles as the
            # right pad of row y (positions are linearised, so x + kx simply runs into the next row's pad)
            W = Wfull
            Wp = W + KW // 2
        else:
            # strip of W tap-grid columns with real neighbour columns on both sides
            W = (Wfull + strips - 1) // strips
            Wp = W + 2 * (KW // 2)
        maxshift = (KH - 1) * Wp + (KW - 1)
        ctot = self._virtual_cin()
        ncn = self.cout_pad // self.N
        is_1x1 = self.kind == "conv" and KD == KH == KW == 1
        btile = lambda kc: self.N * kc * 2

        def fit(KC, ZT, PT, want_slots, min_slots):
            """-> (S_pad, NSLOT, NBST, TPS) or None"""
            CH = KC // 8
            S = 128 * PT + maxshift
            rem = {8: 1, 4: 2, 2: 4}[CH]
            S_pad = S + ((rem - S) % 8)
            slot = CH * S_pad * 16
            tpk = {"conv": KH * KW, "down144": 4, "up144": 4, "unshuffle": 1}[self.kind]  # taps per kz group
            stage_cap = int(os.environ.get("WDNO_BSTAGE", "49152"))  # several taps per issue block (fewer, larger MMA bursts)
            divs = [d for d in range(tpk, 0, -1) if tpk % d == 0 and (d * btile(KC) <= stage_cap or d == 1)]
            # 2-D layers are weight-stream heavy: keep >= 3 weight stages in flight before spending shared memory on
            # slab slots beyond the minimum; 3-D layers (tuned on the smoke U-Net): slots first
            top = min(12, want_slots)
            if not two_d and os.environ.get("WDNO_FIT_ORDER", "taps") == "taps":
                # 3-D layers: taps per weight stage first, slab slots second.  Every stage costs the issuing thread a
                # tcgen05.commit (~67 issue cycles, tools/micro/mma_commit_cost.cu) and an mbarrier round trip, and a stage of one
                # tap is only ZT*PT*KS MMAs long; 256 -> 256 at 10 x 10: TPS 1 / 6 slots 179.3 us, TPS 3 / 4 slots 163.4 us
                # (tools/sweep_generic.py).  Two stages in flight are enough for the L2-resident weight stream.
                for tps in divs:
                    for nslot in range(top, min_slots - 1, -1):
                        for nbst in ((2,) if tps > 1 else (4, 3, 2)):
                            if _BAR_BYTES + _round_up(nslot * slot, 128) + nbst * tps * btile(KC) <= _SMEM_LIMIT:
                                return S_pad, nslot, nbst, tps
                return None
            for min_b, low in (((3, min(top, min_slots + 1)), (2, min_slots)) if two_d else ((2, min_slots),)):
                for nslot in range(top, low - 1, -1):
                    for tps in divs:
                        for nbst in (4, 3, 2):
                            if nbst >= min_b and _BAR_BYTES + _round_up(nslot * slot, 128) + nbst * tps * btile(KC) <= _SMEM_LIMIT:
                                return S_pad, nslot, nbst, tps
            return None

        kcs = [self.kc_override] if self.kc_override else [64, 32, 16]
        if os.environ.get("WDNO_KC"):  # tuning knobs (tools/sweep_tapgemm.py)
            kcs = [int(os.environ["WDNO_KC"])]
        kcs = [k for k in kcs if ctot % k == 0 and not any(c % k for c in self.src_channels[:-1])]
        plan = None
        if is_1x1 and ncn > 1:
            # A-stationary: all K-sets of a work item stay in the slab ring while every N-chunk is computed
            for KC in kcs:
                sets = ctot // KC
                for ZT in [z for z in (4, 2, 1) if z <= D]:
                    need = sets * ZT
                    if need > 12:
                        continue
                    f = fit(KC, ZT, 1, min(12, need + ZT), need)
                    if f:
                        plan = (KC, ZT, 1, 1) + f
                        break
                if plan:
                    break
        zstack = 0
        if (plan is None and self.kind == "conv" and KD in (3, 7) and self.N == 64 and D >= 4 and not self.up2
                and int(os.environ.get("WDNO_ZSTACK", "1"))):
            # kz-stacked MMAs (N = 64..256 per input plane): all P planes of a K-set resident, >= 3 weight stages
            ZT, PT = 4, 1
            P = ZT + KD - 1
            T = KH * KW
            best = None
            for KC in kcs:
                CH = KC // 8
                S = 128 + maxshift
                S_pad = S + (({8: 1, 4: 2, 2: 4}[CH] - S) % 8)
                slot = CH * S_pad * 16
                tile = KD * self.N * KC * 2
                stage_cap = int(os.environ.get("WDNO_ZSTAGE", "40960"))
                for tps in [d for d in range(T, 0, -1) if T % d == 0 and (d * tile <= stage_cap or d == 1)][:2]:
                    for nbst in (4, 3, 2):
                        room = _SMEM_LIMIT - _BAR_BYTES - nbst * tps * tile
                        nslot = min(12, 2 * P, room // slot if room > 0 else 0)
                        while nslot >= P and _BAR_BYTES + _round_up(nslot * slot, 128) + nbst * tps * tile > _SMEM_LIMIT:
                            nslot -= 1
                        if nslot < P:
                            continue
                        # >= 3 taps of weights in flight, then slabs ahead of the MMAs, then fewer / larger issue blocks
                        score = (min(nbst * tps, 6) >= 3 and nslot > P, KC, 1.5 * min(nslot - P, 4) + min(nbst * tps, 6) + 0.5 * tps)
                        if best is None or score > best[0]:
                            best = (score, (KC, ZT, PT, 0, S_pad, nslot, nbst, tps))
            if best is not None:
                plan = best[1]
                zstack = 1
        if plan is None:
            # accumulator shape: pick the plane-group depth ZT whose work-item count balances best over the persistent
            # CTAs (estimated makespan = waves x (MMA cycles + per-item operand/pipeline overhead)); PT = 4 only for 2-D
            sms_ = (torch.cuda.get_device_properties(self.device).multi_processor_count
                    if self.device.type == "cuda" else 148)
            zts = [z for z in (4, 2, 1) if z <= D] or [1]
            if os.environ.get("WDNO_ZT"):
                zts = [int(os.environ["WDNO_ZT"])]
            ntaps = {"conv": KD * KH * KW, "down144": 16, "up144": 4, "unshuffle": 4}[self.kind]
            mma_cyc = 48 if self.N <= 64 else self.N // 2  # N <= 64 is shared-memory-bandwidth bound (tools/micro/mma_rate)
            best = None
            for ZT, PT in [(z, pt) for z in zts for pt in ((4, 1) if (z == 1 and D == 1) else (1,))]:
                # D == 1: four position tiles per item (PT = 4) unless the plane is so small that they would mostly be
                # padding (8x8 planes: 72 positions) -- the estimate below decides
                if self.N > 64:
                    while ZT * PT * 128 > 512:
                        PT = max(1, PT - 1)
                P = ZT + KD - 1
                for KC in kcs:
                    f = fit(KC, ZT, PT, P + int(os.environ.get("WDNO_SLOT_EXTRA", "2")), P)
                    if not f:
                        continue
                    ptiles_ = (H * Wp + 128 * PT - 1) // (128 * PT)
                    items = B * strips * ((D + ZT - 1) // ZT) * ptiles_ * ncn
                    waves = (items + sms_ - 1) // sms_
                    mma_part = ZT * PT * ntaps * (ctot // 16) * mma_cyc
                    if two_d:
                        # weight tiles of one item stream from L2 at ~10 B/cycle/SM (measured on the 8x8 Burgers layers)
                        mma_part = max(mma_part, ntaps * ctot * self.N * 2 // 10)
                    if two_d or os.environ.get("WDNO_FIT_ORDER", "taps") != "taps":
                        per_item = mma_part + 3000 + P * (ctot // KC) * 800 * PT
                    else:
                        # 3-D layers, re-fitted in round 2 (tools/sweep_generic.py): a weight stage costs the issuing thread ~300
                        # cycles (commit + mbarrier round trip + the burst it cannot overlap), a plane of a K-set ~200
                        stages = ntaps * (ctot // KC) // f[3]
                        per_item = mma_part + 3000 + P * (ctot // KC) * 200 * PT + stages * 300
                    est = waves * per_item
                    score = (-est, min(f[1] - P, 2), KC)  # then prefer a full ring (P+2 slots) and the larger KC
                    if best is None or score > best[0]:
                        best = (score, (KC, ZT, PT, 0) + f)
            if best is None:
                return None
            plan = best[1]
            self._last_est = -best[0][0]
        KC, ZT, PT, reuse, S_pad, NSLOT, NBST, TPS = plan
        pk = self._pack(KC, bool(zstack))
        if Wp not in pk["taps_dev"]:
            tl = pk["taps_kyx"]
            taps_c = (Tap * len(tl))(*[Tap(kz, ky * Wp + kx) for (kz, ky, kx) in tl])
            pk["taps_dev"][Wp] = _device_bytes(taps_c, self.device)
        taps_dev = pk["taps_dev"][Wp]
        positions = H * Wp
        ptiles = (positions + 128 * PT - 1) // (128 * PT)
        zgroups = (D + ZT - 1) // ZT
        n_work = B * strips * zgroups * ptiles * (1 if reuse else pk["n_chunks"])
        sms = (torch.cuda.get_device_properties(self.device).multi_processor_count
               if self.device.type == "cuda" else 148)
        p = TapGemmParams()
        p.src_mode = {"conv": 2 if self.up2 else 0, "down144": 1, "unshuffle": 1, "up144": 0}[self.kind]
        p.B, p.D, p.H, p.W = B, D, H, W
        p.strips, p.Wfull = strips, Wfull
        p.fold = 1 if fold else 0
        p.KD, p.pz, p.py, p.px = KD, KD // 2, KH // 2, KW // 2
        p.Wp, p.maxshift = Wp, maxshift
        p.ZT, p.PT, p.KC, p.N, p.n_chunks = ZT, PT, KC, self.N, pk["n_chunks"]
        p.chunks = pk["chunks"].data_ptr()
        p.sets = pk["sets"].data_ptr()
        p.taps = taps_dev.data_ptr()
        p.wpacked = pk["wpacked"].data_ptr()
        p.NSLOT, p.NBST, p.S_pad = NSLOT, NBST, S_pad
        p.TPS, p.reuse = TPS, reuse
        p.n_taps = len(pk["taps_kyx"])
        p.n_sets = pk["n_sets"]
        p.zstack = zstack
        p.grid = max(1, min(n_work, sms))
        p.cluster = 0
        # WDNO_CLUSTER=1 (opt-in): CTA pairs (clusters of 2) share the weight stream -- every tile is fetched from L2 once per pair by
        # a multicast copy.  Built to test whether the layers that need 15-16 B/clk/SM of weight tiles are held back by the L2
        # (every SM pulls the same lines); measured neutral on the 256 -> 256 and 64 -> 64 layers, so it stays off.
        units = n_work // (1 if reuse else pk["n_chunks"])
        if (os.environ.get("WDNO_CLUSTER", "0") == "1" and not reuse and not fold and units % 2 == 0 and n_work >= 4
                and self.device.type == "cuda"):
            smem = self._smem_estimate(KC, S_pad, NSLOT, NBST, TPS, zstack)
            cap = _cluster_cta_cap(smem)
            grid = min(n_work, cap) & ~1
            if grid >= 2 and grid >= (p.grid * 9) // 10:       # keep (almost) every SM busy
                p.grid = grid
                p.cluster = 2
        return p

    def _smem_estimate(self, KC, S_pad, NSLOT, NBST, TPS, zstack):
        slot = (KC // 8) * S_pad * 16
        tile = self.N * KC * 2 * (self.KD if zstack else 1)
        return _BAR_BYTES + _round_up(NSLOT * slot, 128) + NBST * TPS * tile

    # ------------------------------------------------------------------ launch
    def __call__(self, src0, src1=None, *, coef0=None, coef1=None, out=None, resid=None, stats=None,
                 groups=8, out_fp32_bfchw=False):
        """src*: fp16 [B, D, Hs, Ws, C] contiguous.  coef*: (a, c) fp32 [B, C] pairs -> silu(a*x+c) fused on load.
        Returns the output tensor (fp16 [B, D, Ho, Wo, Cout], or fp32 [B, D, Cout, H, W] if out_fp32_bfchw)."""
        L = _lib.lib()
        assert src0.dtype == torch.float16 and src0.is_contiguous() and src0.dim() == 5
        B, D, Hs, Ws, C0 = src0.shape
        if self.kind in ("down144", "unshuffle"):
            H, W = Hs // 2, Ws // 2
        elif self.kind == "conv" and self.up2:
            H, W = Hs * 2, Ws * 2
        else:
            H, W = Hs, Ws
        if self._c1 is not None and coef0 is None and coef1 is None and stats is None:
            return self._call_1x1(L, src0, src1, out, resid, out_fp32_bfchw)
        p0 = self._plan(B, D, H, W)   # may fold the batch of a 2-D layer into depth planes (p.B * p.D == B * D)
        p = TapGemmParams.from_buffer_copy(p0)
        assert C0 == self.src_channels[0], (C0, self.src_channels)
        p.src[0] = src0.data_ptr()
        p.src_c[0] = C0
        if src1 is not None:
            assert src1.dtype == torch.float16 and src1.is_contiguous() and src1.shape[:4] == src0.shape[:4]
            assert src1.shape[4] == self.src_channels[1]
            p.src[1] = src1.data_ptr()
            p.src_c[1] = src1.shape[4]
        else:
            assert len(self.src_channels) == 1
        for i, cf in enumerate((coef0, coef1)):
            if cf is not None:
                a, c = cf
                assert a.dtype == torch.float32 and c.dtype == torch.float32 and a.is_contiguous() and c.is_contiguous()
                assert a.shape == (B, self.src_channels[i]) and c.shape == a.shape
                p.coef_a[i] = a.data_ptr()
                p.coef_c[i] = c.data_ptr()
        Ho, Wo = (2 * H, 2 * W) if self.kind == "up144" else (H, W)
        if out_fp32_bfchw:
            assert self.kind == "conv"
            if out is None:
                out = torch.empty((B, D, self.cout, H, W), dtype=torch.float32, device=src0.device)
            assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (B, D, self.cout, H, W)
            p.out_mode, p.out_c = 2, self.cout
        else:
            if out is None:
                out = torch.empty((B, D, Ho, Wo, self.cout), dtype=torch.float16, device=src0.device)
            assert out.dtype == torch.float16 and out.is_contiguous() and out.shape == (B, D, Ho, Wo, self.cout)
            assert self.cout % 8 == 0
            p.out_mode, p.out_c = (1 if self.kind == "up144" else 0), self.cout
        p.out = out.data_ptr()
        if self.bias is not None:
            p.bias = self.bias.data_ptr()
            p.bias_len = self.bias.numel()
        if resid is not None:
            assert resid.dtype == torch.float16 and resid.is_contiguous() and resid.shape == out.shape
            p.resid = resid.data_ptr()
        if stats is not None:
            assert stats.dtype == torch.float64 and stats.is_contiguous() and stats.numel() == B * groups * 2
            assert self.cout % groups == 0 and (self.cout // groups) % 8 == 0
            p.stats = stats.data_ptr()
            p.G, p.cpg = groups, self.cout // groups
        tm = TapGemm.timing
        flops = 0.0
        if tm is not None or _timing.sink is not None:
            # algorithmic FLOPs of the reference layer (identity K-sets that carry a fused residual are not counted)
            flops = 2.0 * B * D * H * W * self.cout * getattr(self, "algo_cin", self.cin) * self.w.shape[2] * self.w.shape[3] * self.w.shape[4]
            if self.kind == "unshuffle":
                flops *= 4
        meta = (self.kind, self.cin, self.cout, self.KD, self.KH, self.KW, B, D, H, W)
        if tm is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        with _timing.span("tapgemm", flops=flops, meta=meta):
            _lib.check(L.wdno_tapgemm(C.byref(p), _lib.current_stream_ptr()), "tapgemm")
        if tm is not None:
            e1.record()
            tm.append((e0, e1, flops, meta))
        return out

    def _call_1x1(self, L, src0, src1, out, resid, out_fp32_bfchw):
        B, D, H, W, C0 = src0.shape
        c1 = self._c1
        assert C0 == self.src_channels[0], (C0, self.src_channels)
        if src1 is not None:
            assert src1.dtype == torch.float16 and src1.is_contiguous() and src1.shape[:4] == src0.shape[:4]
            assert src1.shape[4] == self.src_channels[1]
        else:
            assert len(self.src_channels) == 1
        if out_fp32_bfchw:
            assert resid is None
            if out is None:
                out = torch.empty((B, D, self.cout, H, W), dtype=torch.float32, device=src0.device)
            assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (B, D, self.cout, H, W)
        else:
            if out is None:
                out = torch.empty((B, D, H, W, self.cout), dtype=torch.float16, device=src0.device)
            assert out.dtype == torch.float16 and out.is_contiguous() and out.shape == (B, D, H, W, self.cout)
            if resid is not None:
                assert resid.dtype == torch.float16 and resid.is_contiguous() and resid.shape == out.shape
        tm = TapGemm.timing
        if tm is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        npos = B * D * H * W
        flops = 2.0 * npos * self.cout * getattr(self, "algo_cin", self.cin)
        # algorithmic bytes: fp16 sources in, output (fp16, or fp32 eps) out, fp16 residual in
        nbytes = npos * (2.0 * sum(self.src_channels) + (4.0 if out_fp32_bfchw else 2.0) * self.cout
                         + (2.0 * self.cout if resid is not None else 0.0))
        meta = ("conv1x1", self.cin, self.cout, 1, 1, 1, B, D, H, W)
        with _timing.span("conv1x1", flops=flops, bytes=nbytes, meta=meta):
            _lib.check(L.wdno_conv1x1(src0.data_ptr(), C0, src1.data_ptr() if src1 is not None else None,
                                      src1.shape[4] if src1 is not None else 0, c1["w"].data_ptr(), c1["npad"],
                                      c1["bias"].data_ptr() if c1["bias"] is not None else None,
                                      resid.data_ptr() if resid is not None else None, out.data_ptr(), npos, self.cout,
                                      2 if out_fp32_bfchw else 0, H * W, _lib.current_stream_ptr()), "conv1x1")
        if tm is not None:
            e1.record()
            tm.append((e0, e1, flops, meta))
        return out

    # bench.py sets this to a list to collect (start event, end event, algorithmic FLOPs, shape) per launch
    timing = None
