"""Training step on the engine: a differentiable `model(x, t)` whose backward runs in
libwdno_b200.so.  Reference: `Trainer.train` -> `loss = self.model(state)`; `accelerator.backward(loss)`; clip 1.0; Adam; EMA
(smoke/ddpm/diffusion_2d.py:1257-1307, burgers/ddpm_burgers/train_diffusion.py:187-237), where every gradient comes from
torch autograd over cuDNN / cuBLAS.

What runs where (this file is host logic only):
  * forward: the inference kernels (tap-GEMM with GroupNorm statistics in the epilogue and GroupNorm-apply + SiLU fused into
    the next operand load, fused attention blocks), keeping the pre-norm convolution outputs y, the per-(sample, channel)
    affine (a, c) and the block inputs;
  * backward, convolutions (>= 93 % of the FLOPs): dgrad = the SAME tap-GEMM kernel on transposed / flipped weight tiles
    (stride-2 conv <-> transposed conv swap kinds), wgrad = `wdno_wgrad` (csrc/wgrad.cu), GroupNorm + (scale, shift) + SiLU
    backward = `wdno_gn_bwd_reduce / _finalize / _apply` (csrc/train.cu);
  * backward, attention blocks and the time-embedding MLP (6 % / < 0.01 % of the FLOPs): recomputed in fp32 torch ops under
    autograd, block by block (interim: their CUDA backward kernels are the next step; stated in DESIGN.md);
  * activation gradients are fp16 channels-last in a scaled domain (loss gradient x 2^k, undone in the fp32 parameter-gradient
    epilogues), parameter gradients accumulate in fp32 into ONE flat buffer (p.grad are views), so clipping, Adam, EMA and the
    data-parallel all-reduce are single launches / collectives over flat memory.
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, ops
from .tapgemm import TapGemm


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class WgradGroup(C.Structure):
    _fields_ = [("dz", C.c_int32), ("dy", C.c_int32), ("dx_min", C.c_int32), ("span", C.c_int32), ("n", C.c_int32),
                ("dxo", C.c_int32 * 3), ("out", C.c_int64 * 3)]


class WgradParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p), ("dbias", C.c_void_p),
                ("B", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("Hs", C.c_int32), ("Ws", C.c_int32), ("Cx", C.c_int32),
                ("sy", C.c_int32), ("sx", C.c_int32), ("ph_y", C.c_int32), ("ph_x", C.c_int32),
                ("Cy", C.c_int32), ("m_valid", C.c_int32), ("cx_off", C.c_int32), ("cx_n", C.c_int32),
                ("n_total", C.c_int32), ("n_off", C.c_int32), ("t_total", C.c_int32), ("padw", C.c_int32),
                ("n_groups", C.c_int32), ("groups", C.c_void_p), ("split", C.c_int32), ("scale", C.c_float)]


_bound = False


def _bind():
    global _bound
    if not _bound:
        L = _lib.lib()
        L.wdno_wgrad.restype = C.c_int
        L.wdno_wgrad.argtypes = [C.POINTER(WgradParams), C.c_void_p]
        _bound = True
    return _lib.lib()


def _groups_dev(groups, device):
    arr = (WgradGroup * len(groups))()
    for i, g in enumerate(groups):
        a = arr[i]
        a.dz, a.dy, a.dx_min, a.span, a.n = g["dz"], g["dy"], g["dx_min"], g["span"], len(g["dxo"])
        for j, (d, o) in enumerate(zip(g["dxo"], g["out"])):
            a.dxo[j], a.out[j] = d, o
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().to(device), len(groups)


def _chunk3(items):
    return [items[i:i + 3] for i in range(0, len(items), 3)]


def wgrad_groups_conv(KD, KH, KW):
    """tap groups of a 'same' convolution: taps sharing (kz, ky), kx in runs of <= 3"""
    gs = []
    for kz in range(KD):
        for ky in range(KH):
            for run in _chunk3(list(range(KW))):
                gs.append(dict(dz=kz - KD // 2, dy=ky - KH // 2, dx_min=run[0] - KW // 2, span=run[-1] - run[0],
                               dxo=[k - run[0] for k in run], out=[(kz * KH + ky) * KW + k for k in run]))
    return gs


# (1,4,4) stride (1,2,2) pad (0,1,1): source pixel 2o + k - 1.  Phase a = its parity: a = 0 -> k in {1, 3} at view rows o + {0, 1};
# a = 1 -> k in {0, 2} at view rows o + {-1, 0}
_PH144 = {0: [(1, 0), (3, 1)], 1: [(0, -1), (2, 0)]}


def wgrad_groups_144(a, b):
    gs = []
    for ky, dyv in _PH144[a]:
        ent = _PH144[b]
        dxs = [d for _, d in ent]
        gs.append(dict(dz=0, dy=dyv, dx_min=min(dxs), span=max(dxs) - min(dxs), dxo=[d - min(dxs) for d in dxs],
                       out=[ky * 4 + kx for kx, _ in ent]))
    return gs


class Wgrad:
    """dW (+ db) of one layer.  kind: 'conv' (same padding, incl. 1x1 / Linear), 'down144', 'up144'."""

    def __init__(self, kind, wshape, device):
        self.kind, self.device = kind, torch.device(device)
        if kind == "conv":
            ws = tuple(wshape) + (1,) * (5 - len(wshape)) if len(wshape) == 2 else tuple(wshape)
            if len(ws) == 4:
                ws = ws[:2] + (1,) + ws[2:]
            self.cout, self.cin, self.KD, self.KH, self.KW = ws
            self.t_total = self.KD * self.KH * self.KW
            self.tables = [(1, 1, 0, 0, self.KW // 2) + _groups_dev(wgrad_groups_conv(self.KD, self.KH, self.KW), self.device)]
        elif kind in ("down144", "up144"):
            # down: weight [cout, cin, 1, 4, 4], dY low-res, X high-res.  up (ConvTranspose): weight [cin, cout, 1, 4, 4]; the
            # low-res INPUT plays dY's role and the high-res output gradient plays X's: dW[ci][co][k] = sum x[ci](i) dOut[co](2i+k-1)
            self.cout, self.cin = wshape[0], wshape[1]
            self.t_total = 16
            self.tables = [(2, 2, a, b, 1) + _groups_dev(wgrad_groups_144(a, b), self.device) for a in (0, 1) for b in (0, 1)]
        elif kind == "unshuffle":
            # Rearrange('b c (h p1) (w p2) -> b (c p1 p2) h w') + Conv2d(4c, c', 1) (unet.py:41-45): per phase (p1, p2) a 1x1 layer on
            # the stride-2 view; weight column c*4 + p1*2 + p2 = (input channel c, "tap" p1*2 + p2)
            self.cout, self.cin = wshape[0], wshape[1] // 4
            self.t_total = 4
            self.tables = [(2, 2, a, b, 0) + _groups_dev([dict(dz=0, dy=0, dx_min=0, span=0, dxo=[0], out=[a * 2 + b])], self.device)
                           for a in (0, 1) for b in (0, 1)]
        else:
            raise ValueError(kind)

    def __call__(self, x, dy, dw, dbias, scale, cx_off=0, cx_n=None, n_off=0, n_total=None, m_valid=None):
        """x: fp16 [B,D,Hs,Ws,Cx] (X role), dy: fp16 [B,D,H,W,Cy] (dY role); dw fp32 [Cy, n_total, taps...] += ; dbias fp32 [Cy] +="""
        L = _bind()
        assert x.dtype == torch.float16 and dy.dtype == torch.float16 and x.is_contiguous() and dy.is_contiguous()
        assert dw.dtype == torch.float32 and dw.is_contiguous()
        B, D, H, W, Cy = dy.shape
        _, _, Hs, Ws, Cx = x.shape
        cx_n = Cx - cx_off if cx_n is None else cx_n
        n_total = cx_n if n_total is None else n_total
        m_valid = Cy if m_valid is None else m_valid
        assert dw.numel() == m_valid * n_total * self.t_total, (tuple(dw.shape), m_valid, n_total, self.t_total)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        for (sy, sx, ph_y, ph_x, padw, gdev, ng) in self.tables:
            p = WgradParams()
            p.x, p.dy, p.dw = x.data_ptr(), dy.data_ptr(), dw.data_ptr()
            p.dbias = dbias.data_ptr() if (dbias is not None and ph_y == 0 and ph_x == 0) else None
            p.B, p.D, p.H, p.W, p.Hs, p.Ws, p.Cx = B, D, H, W, Hs, Ws, Cx
            p.sy, p.sx, p.ph_y, p.ph_x = sy, sx, ph_y, ph_x
            p.Cy, p.m_valid, p.cx_off, p.cx_n = Cy, m_valid, cx_off, cx_n
            p.n_total, p.n_off, p.t_total, p.padw = n_total, n_off, self.t_total, padw
            p.n_groups, p.groups = ng, gdev.data_ptr()
            tiles = ((Cy + 63) // 64) * ((cx_n + 63) // 64)
            chunks = B * D * ((H * (W + padw) + 63) // 64)
            p.split = max(1, min(chunks, (6 * sms + ng * tiles - 1) // (ng * tiles)))
            p.scale = scale
            _lib.check(L.wdno_wgrad(C.byref(p), _lib.current_stream_ptr()), "wgrad")


# ---------------------------------------------------------------------------------------------- small wrappers
def gn_bwd(dh, y, a, c, stats, gamma, beta, dgamma, dbeta, groups, count, scale, ss=None, ss_stride=0, d_ss=None,
           dss_stride=0, add=None, eps=1e-5):
    """-> dy (fp16) of z = a*y + c, h = silu(z) with GroupNorm statistics `stats` [B,G,2]; accumulates d_gamma / d_beta and
    writes the (scale | shift) gradient rows"""
    L = _lib.lib()
    B, Cc = y.shape[0], y.shape[-1]
    vox = y.numel() // (B * Cc)
    sums = torch.zeros((B, Cc, 2), dtype=torch.float64, device=y.device)
    _lib.check(L.wdno_gn_bwd_reduce(_p(dh), _p(y), _p(a), _p(c), _p(sums), B, Cc, vox, _lib.current_stream_ptr()), "gn_bwd_reduce")
    k1 = torch.empty((B, groups), dtype=torch.float32, device=y.device)
    k0 = torch.empty_like(k1)
    _lib.check(L.wdno_gn_bwd_finalize(_p(sums), _p(stats), _p(gamma), _p(beta), _p(ss), ss_stride, _p(dgamma), _p(dbeta),
                                      _p(d_ss), dss_stride, _p(k1), _p(k0), B, Cc, groups, float(count), float(eps),
                                      float(scale), _lib.current_stream_ptr()), "gn_bwd_finalize")
    dy = torch.empty_like(y)
    _lib.check(L.wdno_gn_bwd_apply(_p(dh), _p(y), _p(a), _p(c), _p(k1), _p(k0), _p(add), _p(dy), B, Cc, groups, vox,
                                   _lib.current_stream_ptr()), "gn_bwd_apply")
    return dy


def add_f16(a, b):
    out = torch.empty_like(a)
    _lib.check(_lib.lib().wdno_add_f16(_p(a), _p(b), _p(out), a.numel(), _lib.current_stream_ptr()), "add_f16")
    return out


def pack_grad_f16(g, cp, mul):
    B, Fr, Cc, H, W = g.shape
    out = torch.empty((B, Fr, H, W, cp), dtype=torch.float16, device=g.device)
    _lib.check(_lib.lib().wdno_pack_grad_f16(_p(g), _p(out), B, Fr, Cc, H, W, cp, float(mul), _lib.current_stream_ptr()),
               "pack_grad_f16")
    return out


def chan_layernorm_bwd(x, dy, gamma, dgamma, scale, add=None, eps=1e-5):
    Cc = x.shape[-1]
    dx = torch.empty_like(x)
    _lib.check(_lib.lib().wdno_chan_layernorm_bwd(_p(x), _p(dy), _p(gamma), _p(add), _p(dx), _p(dgamma), x.numel() // Cc, Cc,
                                                  float(eps), float(scale), _lib.current_stream_ptr()), "chan_layernorm_bwd")
    return dx


# ---------------------------------------------------------------------------------------------- conv layer with gradients
class ConvLayer:
    """forward plan (shared with the inference engine) + dgrad plans (one per concatenated source) + wgrad of one
    Conv3d / Conv2d / ConvTranspose3d / Linear parameter pair."""

    def __init__(self, fwd, weight, bias, kind, src_channels, need_dgrad=True):
        self.fwd, self.weight, self.bias, self.kind = fwd, weight, bias, kind
        self.src_channels = tuple(src_channels)
        self.up2 = bool(getattr(fwd, "up2", False))   # nearest x2 up-sampling fused into the forward operand load
        dev = fwd.device
        self.wgrad = Wgrad(kind, tuple(weight.shape), dev)
        # stride-1 'same' convolutions with 64 / 128k output channels: the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cu);
        # WDNO_WGRAD_TC=0 keeps every layer on the mma.sync kernel
        self.wgrad_tc = None
        import os
        if kind == "conv" and not self.up2 and os.environ.get("WDNO_WGRAD_TC", "1") != "0":
            w5 = self.wgrad
            if all(WgradTC.supported(w5.cout, min(cs, w5.cin), w5.KD, w5.KH, w5.KW) for cs in self.src_channels):
                self.wgrad_tc = WgradTC(w5.cout, w5.KD, w5.KH, w5.KW, dev)
        self.dgrad = []
        if need_dgrad:
            for w in self._dgrad_weights():
                if kind == "conv":
                    self.dgrad.append(TapGemm(w, None, src_channels=(self._dy_channels(),), device=dev))
                else:
                    self.dgrad.append(TapGemm(w, None, kind={"down144": "up144", "up144": "down144", "unshuffle": "up144"}[kind],
                                              device=dev))

    def _dy_channels(self):
        return (self.weight.shape[0] + 7) // 8 * 8

    def _dgrad_weights(self):
        w = self.weight.detach()
        if self.kind == "unshuffle":
            # dx[2y+p1][2x+p2][c] = sum_co W[co][c*4 + p1*2 + p2] dy[y][x][co]: a transposed (1,4,4) stride-2 convolution whose only
            # non-zero taps are k = p + 1 (each output pixel sees exactly its own source pixel)
            co, c4 = w.shape[0], w.shape[1]
            wt = w.new_zeros((co, c4 // 4, 1, 4, 4))
            wt[:, :, 0, 1:3, 1:3] = w.reshape(co, c4 // 4, 2, 2)
            return [wt]
        if self.kind != "conv":
            return [w]   # the (1,4,4) pair: strided conv <-> transposed conv with the SAME weight tensor
        if w.dim() == 2:
            w = w[:, :, None, None, None]
        elif w.dim() == 4:
            w = w[:, :, None]
        out, off = [], 0
        for cs in self.src_channels:
            real = min(cs, w.shape[1] - off)
            wt = w[:, off:off + real].transpose(0, 1).flip(2, 3, 4)     # [cin_s, cout, taps] : conv of dY with flipped taps
            if real < cs:   # padded source channels (the stem's 42 -> 48): their gradient is never used
                wt = torch.cat((wt, wt.new_zeros((cs - real,) + tuple(wt.shape[1:]))), 0)
            out.append(wt.contiguous())
            off += cs
        return out

    def refresh(self, fwd=True):
        if fwd:
            self.fwd.refresh(self.weight, self.bias)
        for plan, w in zip(self.dgrad, self._dgrad_weights()):
            plan.refresh(w)

    def backward_input(self, dy, s, resid=None):
        """gradient w.r.t. source `s` (fp16 channels-last), optionally + resid (fused into the epilogue)"""
        if self.up2:
            g = self.dgrad[s](dy)            # gradient of the up-sampled operand; its adjoint is the 2x2 sum
            B, D, H2, W2, Cc = g.shape
            out = torch.empty((B, D, H2 // 2, W2 // 2, Cc), dtype=torch.float16, device=g.device)
            _lib.check(_lib.lib().wdno_sumpool2x2_f16(_p(g), _p(out), B * D, H2 // 2, W2 // 2, Cc, _lib.current_stream_ptr()),
                       "sumpool2x2_f16")
            return out if resid is None else add_f16(out, resid)
        return self.dgrad[s](dy, resid=resid)

    def backward_weight(self, srcs, dy, scale):
        """srcs: the forward sources (for 'up144': dy is the INPUT and srcs[0] the output gradient -- see Wgrad)"""
        gw = self.weight.grad
        gb = None if self.bias is None else self.bias.grad
        if self.kind == "up144":
            # weight [cin, cout, 1, 4, 4]: M = cin (the layer input), N = cout (the output gradient); bias grad = column sums of dOut
            self.wgrad(srcs[0], dy, gw, None, scale)
            if gb is not None:
                gb.add_(srcs[0].float().sum(dim=(0, 1, 2, 3)) * scale)
            return
        if self.up2:
            ups = []
            for src in srcs:
                B, D, H, W, Cc = src.shape
                u = torch.empty((B, D, 2 * H, 2 * W, Cc), dtype=torch.float16, device=src.device)
                _lib.check(_lib.lib().wdno_upsample2x_f16(_p(src), _p(u), B * D, H, W, Cc, _lib.current_stream_ptr()), "upsample2x_f16")
                ups.append(u)
            srcs = ups
        off = 0
        n_total = self.weight.shape[1] // (4 if self.kind == "unshuffle" else 1)
        for i, (src, cs) in enumerate(zip(srcs, self.src_channels)):
            real = min(cs, n_total - off)
            if self.wgrad_tc is not None:
                self.wgrad_tc(src, dy, gw, scale, cx_off=0, cx_n=real, n_off=off, n_total=n_total, m_valid=self.weight.shape[0])
                if i == 0 and gb is not None:
                    colsum_f16(dy, gb, scale)
            else:
                self.wgrad(src, dy, gw, gb if i == 0 else None, scale, cx_off=0, cx_n=real, n_off=off, n_total=n_total,
                           m_valid=self.weight.shape[0])
            off += cs


# ---------------------------------------------------------------------------------------------- attention cores, backward
def softmax_attn_bwd(qkv, d_o, n_seq, n_tok, inner, outerT, innerT, tokT, scale, bias=None, rot=None, dbias=None):
    """qkv fp16 [.., 384], d_o fp16 [.., 128] (gradient of the core output) -> dqkv fp16 [.., 384]; dbias fp32 [4,n,n] +="""
    assert qkv.dtype == torch.float16 and d_o.dtype == torch.float16 and qkv.is_contiguous() and d_o.is_contiguous()
    dqkv = torch.empty_like(qkv)
    rc, rs = (None, None) if rot is None else rot
    _lib.check(_lib.lib().wdno_softmax_attn_bwd(_p(qkv), _p(d_o), _p(bias), _p(rc), _p(rs), _p(dqkv), _p(dbias), n_seq, n_tok,
                                                inner, outerT, innerT, tokT, float(scale), _lib.current_stream_ptr()),
               "softmax_attn_bwd")
    return dqkv


_LA_WORK = {}


def linear_attn_bwd(qkv, d_o, n_img, n_pos, scale):
    assert qkv.dtype == torch.float16 and d_o.dtype == torch.float16 and qkv.is_contiguous() and d_o.is_contiguous()
    L = _lib.lib()
    key = (qkv.device, n_img)
    if key not in _LA_WORK:
        _LA_WORK[key] = torch.empty(L.wdno_linear_attn_bwd_work_bytes(n_img), dtype=torch.uint8, device=qkv.device)
    dqkv = torch.empty_like(qkv)
    _lib.check(L.wdno_linear_attn_bwd(_p(qkv), _p(d_o), _p(dqkv), _p(_LA_WORK[key]), n_img, n_pos, float(scale),
                                      _lib.current_stream_ptr()), "linear_attn_bwd")
    return dqkv


class AttnGrad:
    """Backward of Residual(PreNorm(dim, attention)) around one of the three attention cores (conv3d.py:165-184, 232-353):
    recompute LayerNorm -> to_qkv -> core (unfused kernels), then to_out wgrad / dgrad -> core backward -> to_qkv wgrad / dgrad ->
    LayerNorm backward + the residual.  kind: 'temporal' | 'spatial' | 'linear'."""

    def __init__(self, kind, gamma, to_qkv, to_out, device, rel_emb=None, out_norm=None):
        """out_norm: gain of the channel LayerNorm that follows to_out in the Burgers LinearAttention (unet.py:190-199) or None"""
        self.kind, self.gamma, self.rel_emb, self.out_norm = kind, gamma, rel_emb, out_norm
        C_ = to_qkv.weight.shape[1]
        bo = getattr(to_out, "bias", None)
        self.qkv = ConvLayer(TapGemm(to_qkv.weight, None, device=device), to_qkv.weight, None, "conv", (C_,))
        self.out = ConvLayer(TapGemm(to_out.weight, bo, device=device), to_out.weight, bo, "conv", (128,))
        self.own = (self.qkv, self.out)

    def refresh(self):
        for l in self.own:
            l.refresh()

    def backward(self, x, dy, inv, tables=None, scale=32 ** -0.5):
        B, D, H, W, Cc = x.shape
        g = self.gamma.detach().reshape(-1)
        xn = ops.chan_layernorm(x, g)
        qkv = self.qkv.fwd(xn)
        if self.kind == "linear":
            o = ops.linear_attn(qkv, B * D, H * W, scale)
        elif self.kind == "temporal":
            bias, rot = tables
            amap = (B * H * W, D, H * W, D * H * W, 1, H * W)
            o = ops.softmax_attn(qkv, *amap, scale, bias=bias, rot=rot)
        else:
            amap = (B * D, H * W, 1, H * W, 0, 1)
            o = ops.softmax_attn(qkv, *amap, scale)
        dt = dy
        if self.out_norm is not None:
            # y = LayerNorm(to_out(o)) + x: differentiate the output norm first (its input is recomputed)
            t = self.out.fwd(o)
            dt = chan_layernorm_bwd(t, dy, self.out_norm.detach().reshape(-1), self.out_norm.grad, inv)
            del t
        self.out.backward_weight((o,), dt, inv)
        d_o = self.out.backward_input(dt, 0)
        del o
        if self.kind == "linear":
            dqkv = linear_attn_bwd(qkv, d_o, B * D, H * W, scale)
        elif self.kind == "temporal":
            dbias = torch.zeros_like(bias)
            dqkv = softmax_attn_bwd(qkv, d_o, *amap, scale, bias=bias, rot=rot, dbias=dbias)
            # bias[h, i, j] = emb[bucket(i, j), h]  (RelativePositionBias, conv3d.py:74-112)
            self.rel_emb.grad.index_add_(0, self.bucket(D, x.device), dbias.permute(1, 2, 0).reshape(-1, dbias.shape[0]), alpha=inv)
        else:
            dqkv = softmax_attn_bwd(qkv, d_o, *amap, scale)
        del qkv, d_o
        self.qkv.backward_weight((xn,), dqkv, inv)
        dxn = self.qkv.backward_input(dqkv, 0)
        return chan_layernorm_bwd(x, dxn, g, self.gamma.grad, inv, add=dy)

    _buckets = {}

    @classmethod
    def bucket(cls, n, device, num_buckets=32, max_distance=32):
        key = (n, str(device))
        if key not in cls._buckets:
            pos = torch.arange(n)
            k = -(pos[None, :] - pos[:, None])
            half = num_buckets // 2
            ret = (k < 0).long() * half
            k = k.abs()
            max_exact = half // 2
            large = max_exact + (torch.log(k.float() / max_exact) / math.log(max_distance / max_exact) * (half - max_exact)).long()
            large = torch.min(large, torch.full_like(large, half - 1))
            cls._buckets[key] = (ret + torch.where(k < max_exact, k, large)).reshape(-1).to(device)
        return cls._buckets[key]


# ---------------------------------------------------------------------------------------------- wgrad on tcgen05
class WgradTcJob(C.Structure):
    _fields_ = [("kz", C.c_int32), ("n_taps", C.c_int32), ("qmin", C.c_int32), ("span", C.c_int32), ("m0", C.c_int32),
                ("n0", C.c_int32), ("shift", C.c_int32 * 8), ("out_hi", C.c_int64 * 8), ("out_lo", C.c_int64 * 8)]


class WgradTcParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p),
                ("B", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("Cx", C.c_int32), ("Cy", C.c_int32), ("m_valid", C.c_int32),
                ("cx_off", C.c_int32), ("cx_n", C.c_int32), ("n_total", C.c_int32), ("n_off", C.c_int32), ("t_total", C.c_int32),
                ("padw", C.c_int32), ("pz", C.c_int32), ("stack", C.c_int32), ("nx", C.c_int32), ("ncols", C.c_int32),
                ("stages", C.c_int32), ("n_jobs", C.c_int32), ("jobs", C.c_void_p), ("split", C.c_int32), ("scale", C.c_float),
                ("swap_lbo_sbo", C.c_int32), ("reserved0", C.c_int32)]


_tc_bound = False


def _bind_tc():
    global _tc_bound
    L = _lib.lib()
    if not _tc_bound:
        L.wdno_wgrad_tc.restype = C.c_int
        L.wdno_wgrad_tc.argtypes = [C.POINTER(WgradTcParams), C.c_int, C.c_void_p]
        L.wdno_wgrad_tc_smem_bytes.restype = C.c_int64
        L.wdno_wgrad_tc_smem_bytes.argtypes = [C.POINTER(WgradTcParams), C.c_int]
        _tc_bound = True
    return L


class WgradTC:
    """tcgen05 weight gradient of one stride-1 'same' convolution (csrc/wgrad_tc.cu).  `supported()` says whether the layer
    geometry fits (64 or a multiple of 128 output channels; a window of <= 64 or a multiple of 64 input channels)."""

    SWAP = 0   # descriptor-stride convention (settled by tools/probe_wgrad_tc.py)

    @staticmethod
    def supported(cout, cin_window, KD, KH, KW):
        if KD * KH * KW == 1:
            return False                       # 1x1 layers: the mma.sync kernel (K loop too short to fill TMEM with taps)
        if not (cout == 64 and KD > 1) and cout % 128:
            return False
        return cin_window <= 64 or cin_window % 64 == 0

    def __init__(self, cout, KD, KH, KW, device):
        self.cout, self.KD, self.KH, self.KW = cout, KD, KH, KW
        self.device = torch.device(device)
        self.stack = 1 if cout == 64 else 0
        self._jobs = {}

    def _job_table(self, Wp, cx_n):
        key = (Wp, cx_n)
        if key in self._jobs:
            return self._jobs[key]
        KD, KH, KW = self.KD, self.KH, self.KW
        nx = min(64, (cx_n + 15) // 16 * 16)
        per = max(1, min(8, 512 // 64))
        taps = sorted(((ky - KH // 2) * Wp + (kx - KW // 2), ky, kx) for ky in range(KH) for kx in range(KW))
        ngrp = (len(taps) + per - 1) // per
        size = (len(taps) + ngrp - 1) // ngrp
        groups = [taps[i:i + size] for i in range(0, len(taps), size)]
        kzs = list(range(KD - 1, -1, -2)) if self.stack else list(range(KD))
        m_tiles = [0] if self.stack else list(range(0, self.cout, 128))
        n_tiles = list(range(0, cx_n, 64))
        jobs = []
        for kz in kzs:
            for grp in groups:
                for m0 in m_tiles:
                    for n0 in n_tiles:
                        j = WgradTcJob()
                        j.kz, j.n_taps, j.m0, j.n0 = kz, len(grp), m0, n0
                        j.qmin = grp[0][0]
                        j.span = grp[-1][0] - grp[0][0]
                        for i, (sh, ky, kx) in enumerate(grp):
                            j.shift[i] = sh - grp[0][0]
                            j.out_hi[i] = (kz * KH + ky) * KW + kx
                            j.out_lo[i] = ((kz - 1) * KH + ky) * KW + kx if (self.stack and kz >= 1) else -1
                        jobs.append(j)
        arr = (WgradTcJob * len(jobs))(*jobs)
        dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().to(self.device)
        self._jobs[key] = (dev, len(jobs), max(j.span for j in jobs), nx)
        return self._jobs[key]

    def __call__(self, x, dy, dw, scale, cx_off=0, cx_n=None, n_off=0, n_total=None, m_valid=None):
        L = _bind_tc()
        B, D, H, W, Cy = dy.shape
        Cx = x.shape[-1]
        cx_n = Cx - cx_off if cx_n is None else cx_n
        n_total = cx_n if n_total is None else n_total
        padw = self.KW // 2
        jobs_dev, n_jobs, max_span, nx = self._job_table(W + padw, cx_n)
        p = WgradTcParams()
        p.x, p.dy, p.dw = x.data_ptr(), dy.data_ptr(), dw.data_ptr()
        p.B, p.D, p.H, p.W, p.Cx, p.Cy = B, D, H, W, Cx, Cy
        p.m_valid = Cy if m_valid is None else m_valid
        p.cx_off, p.cx_n, p.n_total, p.n_off, p.t_total = cx_off, cx_n, n_total, n_off, self.KD * self.KH * self.KW
        p.padw, p.pz, p.stack, p.nx, p.ncols = padw, self.KD // 2, self.stack, nx, 64
        p.n_jobs, p.jobs = n_jobs, jobs_dev.data_ptr()
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        items = B * (D + (1 if self.stack else 0)) * ((H * (W + padw) + 127) // 128)
        p.split = max(1, min(items, sms // n_jobs if n_jobs <= sms else 1))
        p.scale, p.swap_lbo_sbo = scale, WgradTC.SWAP
        p.stages = 3
        if L.wdno_wgrad_tc_smem_bytes(C.byref(p), max_span) > 227 * 1024:
            p.stages = 2
        _lib.check(L.wdno_wgrad_tc(C.byref(p), max_span, _lib.current_stream_ptr()), "wgrad_tc")


def colsum_f16(dy, out, scale):
    Cc = dy.shape[-1]
    assert out.dtype == torch.float32 and out.numel() >= Cc
    _lib.check(_lib.lib().wdno_colsum_f16(_p(dy), dy.numel() // Cc, Cc, _p(out), float(scale), _lib.current_stream_ptr()), "colsum_f16")
