"""Host-side plans of the fused attention blocks (csrc/attn_fused.cu): weights re-packed once per parameter snapshot
into mma.sync fragment order, plus the small per-call workspace.

  LinAttnBlock   Residual(PreNorm(dim, SpatialLinearAttention(dim)))      conv3d.py:165-184, 232-258, 426-427, 459
  TemporalBlock  Residual(PreNorm(dim, EinopsToAndFrom(Attention(dim))))  conv3d.py:262-353, 383, 397, 428, 460
"""
import ctypes as C

import torch

from . import _lib, _timing


def pack_b_frags(w):
    """w [N, K] (out-features x in-features, N % 8 == 0, K % 32 == 0) -> fp16 [N/8, K/32, 32 lanes, 8]: the B fragments
    (col-major k x n) of two consecutive k-steps of mma.m16n8k16 for one 8-wide n-tile, one 16-byte load per lane."""
    N, K = w.shape
    assert N % 8 == 0 and K % 32 == 0
    r = w.reshape(N // 8, 8, K // 32, 2, 2, 4, 2)          # nt, g, kp, ks, reg(+0/+8), q, half
    return r.permute(0, 2, 1, 5, 3, 4, 6).contiguous().reshape(N // 8, K // 32, 32, 8).to(torch.float16)


def pack_a_frags(w):
    """w [M, K] (M % 16 == 0, K % 16 == 0) -> fp16 [M/16, K/16, 32 lanes, 8]: the A fragment (row-major m x k) of one
    k-step: a0=(g,2q) a1=(g+8,2q) a2=(g,2q+8) a3=(g+8,2q+8)."""
    M, K = w.shape
    assert M % 16 == 0 and K % 16 == 0
    r = w.reshape(M // 16, 2, 8, K // 16, 2, 4, 2)         # mt, rh(+0/+8), g, ks, kh(+0/+8), q, half
    return r.permute(0, 3, 2, 5, 4, 1, 6).contiguous().reshape(M // 16, K // 16, 32, 8).to(torch.float16)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class LinAttnBlock:
    """y = x + to_out(linear_attention(to_qkv(LayerNorm(x))))  on fp16 channels-last [B, D, H, W, C] (images = B*D)."""

    def __init__(self, gamma, w_qkv, w_out, b_out, heads=4, dim_head=32, device="cuda"):
        assert heads == 4 and dim_head == 32
        dev = torch.device(device)
        self._src = (gamma, w_qkv, w_out, b_out)
        self.C = w_qkv.shape[1]
        assert w_qkv.shape[0] == 384 and self.C in (64, 128, 256)
        self.gamma, self.wq, self.wkv, self.wout, self.bias = (t if t is None else t.to(dev) for t in self._packed())
        assert self.wout.shape == (self.C, 128)
        self.scale = dim_head ** -0.5
        self._work = {}
        # C = 64 / 128: every product on tcgen05 (csrc/linattn_tc.cu); WDNO_LINATTN_TC=0 keeps the mma.sync kernels
        import os
        self.tc = self.C in (64, 128) and os.environ.get("WDNO_LINATTN_TC", "1") != "0"
        self._tc_force = os.environ.get("WDNO_LINATTN_TC", "1") == "force"
        self._tc_ok = {}
        if self.tc:
            self.wq_c, self.wkv_c = (t.to(dev) for t in self._packed_canon())

    def _packed_canon(self):
        """UMMA canonical K-major operands [C/8][rows][8]: W_q (128 rows) and W_k | W_v (256 rows), LayerNorm gain folded in"""
        gamma, w_qkv = self._src[0], self._src[1]
        wq = w_qkv.detach().float().reshape(w_qkv.shape[0], -1) * gamma.detach().float().reshape(1, -1).to(w_qkv.device)   # [384, C]
        canon = lambda w: w.reshape(w.shape[0], w.shape[1] // 8, 8).permute(1, 0, 2).contiguous().to(torch.float16)
        return canon(wq[:128]), canon(wq[128:])

    def _packed(self):
        gamma, w_qkv, w_out, b_out = self._src
        wq = w_qkv.detach().float().reshape(w_qkv.shape[0], -1)    # [384, C]
        # k / v: per (section, head) a [32, C] A operand
        kv = wq[128:].reshape(2, 4, 32, self.C)
        wkv = torch.stack([torch.stack([pack_a_frags(kv[s, h]) for h in range(4)]) for s in range(2)]).contiguous()
        return (gamma.detach().float().reshape(-1).contiguous(), pack_b_frags(wq[:128]), wkv,
                w_out.detach().float().reshape(w_out.shape[0], -1).contiguous(),   # [C, 128]
                None if b_out is None else b_out.detach().float().contiguous())

    def refresh(self):
        """re-pack from the live parameters into the existing device buffers (addresses stay valid for captured graphs)"""
        for dst, src in zip((self.gamma, self.wq, self.wkv, self.wout, self.bias), self._packed()):
            if dst is not None and dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        if self.tc:
            for dst, src in zip((self.wq_c, self.wkv_c), self._packed_canon()):
                dst.copy_(src)

    def __call__(self, x, eps=1e-5):
        assert x.dtype == torch.float16 and x.is_contiguous() and x.dim() == 5 and x.shape[-1] == self.C
        B, D, H, W, _ = x.shape
        n_img, n_pos = B * D, H * W
        L = _lib.lib()
        key = (n_img, n_pos)
        if key not in self._work:
            nbytes = L.wdno_linattn_work_bytes(n_img, n_pos, self.C)
            assert nbytes > 0
            self._work[key] = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        y = torch.empty_like(x)
        # algorithmic work of the block (conv3d.py:232-258): one read + one write of the fp16 residual stream; qkv and output
        # projections + the two 32x32-per-head context products
        ntok = n_img * n_pos
        with _timing.span("linattn_block", flops=2.0 * ntok * (self.C * 384 + 128 * self.C + 2 * 128 * 32),
                          bytes=4.0 * ntok * self.C, meta=(self.C, n_img, n_pos)):
            self._launch(L, x, y, key, n_img, n_pos, eps)
        return y

    def _launch(self, L, x, y, key, n_img, n_pos, eps):
        if self.tc:
            if key not in self._tc_ok:
                # one persistent CTA walks whole images: with fewer images than half the SMs the mma.sync kernels (which split
                # an image over many blocks) fill the machine better (6 images: 69 vs 47 us); WDNO_LINATTN_TC=force overrides
                enough = n_img >= 74 or self._tc_force
                self._tc_ok[key] = enough and L.wdno_linattn_tc_supported(n_img, n_pos, self.C) == 1
            if self._tc_ok[key]:
                _lib.check(L.wdno_linattn_block_tc(_p(x), _p(y), _p(self.wq_c), _p(self.wkv_c), _p(self.wout),
                                                   _p(self.bias), _p(self._work[key]), n_img, n_pos, self.C, self.scale,
                                                   float(eps), _lib.current_stream_ptr()), "linattn_block_tc")
                return
        _lib.check(L.wdno_linattn_block(_p(x), _p(y), _p(self.gamma), _p(self.wq), _p(self.wkv), _p(self.wout), _p(self.bias),
                                        _p(self._work[key]), n_img, n_pos, self.C, self.scale, float(eps),
                                        _lib.current_stream_ptr()), "linattn_block")


class TemporalBlock:
    """y = x + to_out(softmax(q k^T + rel_pos_bias) v)  along the frame axis of fp16 channels-last [B, D, H, W, C],
    q/k rotary-embedded; (q, k, v) = to_qkv(LayerNorm(x)).  D <= 32, H*W even."""

    def __init__(self, gamma, w_qkv, w_out, heads=4, dim_head=32, device="cuda"):
        assert heads == 4 and dim_head == 32
        dev = torch.device(device)
        self._src = (gamma, w_qkv, w_out)
        self.C = w_qkv.shape[1]
        assert w_qkv.shape[0] == 384 and self.C in (64, 128, 256) and tuple(w_out.shape[:2]) == (self.C, 128)
        self.gamma, self.wqk, self.wv, self.wo = (t.to(dev) for t in self._packed())
        self.scale = dim_head ** -0.5
        # C = 64 (the full-resolution instances): WDNO_TATTN_TC=2 (default) every product on tcgen05, softmax per accumulator row
        # (csrc/tattn_row.cu); =1 projections on tcgen05 + mma.sync attention (csrc/tattn_tc.cu); =0 the mma.sync kernel
        import os
        mode = os.environ.get("WDNO_TATTN_TC", "2")
        self.tc = self.C == 64 and mode == "1"
        self.row = self.C == 64 and mode not in ("0", "1")
        if self.tc:
            self.wqkv_c, self.wo_c = (t.to(dev) for t in self._packed_canon())
        if self.row:
            self.wqkv_g, self.wo_g = (t.to(dev) for t in self._packed_canon(fold_gamma=True))

    def _packed_canon(self, fold_gamma=False):
        """UMMA canonical K-major operands: [K/8][rows][8]; fold_gamma: W_qkv diag(gamma) (the row kernel normalises without gain)"""
        gamma, w_qkv, w_out = self._src
        wq = w_qkv.detach().float().reshape(w_qkv.shape[0], -1)    # [384, C]
        if fold_gamma:
            # the row kernel also wants the attention scale and log2(e) (its softmax runs in base 2) in the q rows
            wq = wq * gamma.detach().float().reshape(1, -1).to(wq.device)
            wq = torch.cat([wq[:128] * (self.scale * 1.4426950408889634), wq[128:]])
        wo = w_out.detach().float().reshape(w_out.shape[0], -1)    # [C, 128]
        canon = lambda w: w.reshape(w.shape[0], w.shape[1] // 8, 8).permute(1, 0, 2).contiguous().to(torch.float16)
        return canon(wq), canon(wo)

    def _packed(self):
        gamma, w_qkv, w_out = self._src
        wq = w_qkv.detach().float().reshape(w_qkv.shape[0], -1)    # [384, C]
        wo = w_out.detach().float().reshape(w_out.shape[0], -1)    # [C, 128]
        return (gamma.detach().float().reshape(-1).contiguous(), pack_b_frags(wq[:256]),          # q | k as B operands
                torch.stack([pack_a_frags(wq[256 + 32 * h: 288 + 32 * h]) for h in range(4)]).contiguous(), pack_b_frags(wo))

    def refresh(self):
        for dst, src in zip((self.gamma, self.wqk, self.wv, self.wo), self._packed()):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        if self.tc:
            for dst, src in zip((self.wqkv_c, self.wo_c), self._packed_canon()):
                dst.copy_(src)
        if self.row:
            for dst, src in zip((self.wqkv_g, self.wo_g), self._packed_canon(fold_gamma=True)):
                dst.copy_(src)

    def __call__(self, x, bias=None, rot=None, eps=1e-5):
        """bias fp32 [4, D, D] or None; rot = (cos, sin) fp32 [D, 16] or None."""
        assert x.dtype == torch.float16 and x.is_contiguous() and x.dim() == 5 and x.shape[-1] == self.C
        B, D, H, W, _ = x.shape
        y = torch.empty_like(x)
        rc, rs = (rot if rot is not None else (None, None))
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.is_contiguous() and tuple(bias.shape) == (4, D, D)
        if rc is not None:
            assert rc.dtype == torch.float32 and rc.is_contiguous() and tuple(rc.shape) == (D, 16) and tuple(rs.shape) == (D, 16)
        ntok = B * D * H * W
        # algorithmic work (conv3d.py:262-353): one read + one write of the fp16 residual stream; projections + QK^T + PV
        with _timing.span("tattn_block", flops=2.0 * ntok * (self.C * 384 + 128 * self.C + 2 * 128 * D),
                          bytes=4.0 * ntok * self.C, meta=(self.C, B, D, H * W)):
            if self.row:
                _lib.check(_lib.lib().wdno_tattn_block_row(_p(x), _p(y), _p(self.wqkv_g), _p(self.wo_g), _p(bias), _p(rc), _p(rs),
                                                           B, D, H * W, self.C, self.scale, float(eps),
                                                           _lib.current_stream_ptr()), "tattn_block_row")
                return y
            if self.tc:
                _lib.check(_lib.lib().wdno_tattn_block_tc(_p(x), _p(y), _p(self.gamma), _p(self.wqkv_c), _p(self.wo_c), _p(bias),
                                                          _p(rc), _p(rs), B, D, H * W, self.C, self.scale, float(eps),
                                                          _lib.current_stream_ptr()), "tattn_block_tc")
                return y
            _lib.check(_lib.lib().wdno_tattn_block(_p(x), _p(y), _p(self.gamma), _p(self.wqk), _p(self.wv), _p(self.wo), _p(bias),
                                                   _p(rc), _p(rs), B, D, H * W, self.C, self.scale, float(eps),
                                                   _lib.current_stream_ptr()), "tattn_block")
        return y
