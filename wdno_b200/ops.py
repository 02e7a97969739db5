"""Thin torch-tensor wrappers over the C ABI (include/wdno_b200.h).  Every function enqueues on the
current torch CUDA stream and raises if libwdno_b200.so is missing -- there is no fallback path."""
import ctypes as C

import torch

from . import _lib
from ._abi import MAX_COND_OPS, CondOp


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), f"{name}: need contiguous CUDA {dtype}"


def _st():
    return _lib.current_stream_ptr()


def pack_bfchw_f16(x, cp):
    """fp32 [B,F,C,H,W] -> fp16 [B,F,H,W,cp]"""
    _chk(x, torch.float32, "x")
    B, F, Cc, H, W = x.shape
    out = torch.empty((B, F, H, W, cp), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().wdno_pack_bfchw_f16(_p(x), _p(out), B, F, Cc, H, W, cp, _st()), "pack_bfchw_f16")
    return out


def gn_finalize(stats, gamma, beta, ss, ss_off, ss_stride, B, Cc, G, count, eps=1e-5):
    """-> (a, c) fp32 [B, C].  ss: fp32 [B, ss_stride] holding (scale|shift) of this block at column ss_off, or None."""
    a = torch.empty((B, Cc), dtype=torch.float32, device=stats.device)
    c = torch.empty_like(a)
    ssp = None if ss is None else C.c_void_p(ss.data_ptr() + 4 * ss_off)
    _lib.check(_lib.lib().wdno_gn_finalize(_p(stats), _p(gamma), _p(beta), ssp, ss_stride, _p(a), _p(c), B, Cc, G,
                                          float(count), float(eps), _st()), "gn_finalize")
    return a, c


def gn_silu_add(y, a, c, resid=None):
    _chk(y, torch.float16, "y")
    B, Cc = y.shape[0], y.shape[-1]
    vox = y.numel() // (B * Cc)
    out = torch.empty_like(y)
    _lib.check(_lib.lib().wdno_gn_silu_add(_p(y), _p(a), _p(c), _p(resid), _p(out), B, Cc, vox, _st()), "gn_silu_add")
    return out


def chan_layernorm(x, gamma, eps=1e-5, resid=None):
    _chk(x, torch.float16, "x")
    Cc = x.shape[-1]
    out = torch.empty_like(x)
    _lib.check(_lib.lib().wdno_chan_layernorm(_p(x), _p(gamma), _p(resid), _p(out), x.numel() // Cc, Cc, float(eps), _st()),
               "chan_layernorm")
    return out


def time_mlp(time_f32, w1, b1, w2, b2, theta=10000.0):
    B = time_f32.shape[0]
    tdim, dim = w1.shape
    emb = torch.empty((B, tdim), dtype=torch.float32, device=time_f32.device)
    emb_silu = torch.empty_like(emb)
    _lib.check(_lib.lib().wdno_time_mlp(_p(time_f32), _p(w1), _p(b1), _p(w2), _p(b2), _p(emb), _p(emb_silu), B, dim, tdim,
                                       float(theta), _st()), "time_mlp")
    return emb, emb_silu


def small_linear(x, w, b):
    B, K = x.shape
    J = w.shape[0]
    out = torch.empty((B, J), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().wdno_small_linear(_p(x), _p(w), _p(b), _p(out), B, K, J, _st()), "small_linear")
    return out


def softmax_attn(qkv, n_seq, n_tok, inner, outerT, innerT, tokT, scale, bias=None, rot=None):
    """qkv fp16 [..., 384] -> fp16 [..., 128] (same leading token layout)"""
    _chk(qkv, torch.float16, "qkv")
    assert qkv.shape[-1] == 384
    out = torch.empty(qkv.shape[:-1] + (128,), dtype=torch.float16, device=qkv.device)
    rc, rs = (None, None) if rot is None else rot
    _lib.check(_lib.lib().wdno_softmax_attn(_p(qkv), _p(out), _p(bias), _p(rc), _p(rs), n_seq, n_tok, inner, outerT, innerT,
                                           tokT, float(scale), _st()), "softmax_attn")
    return out


def linear_attn(qkv, n_img, n_pos, scale):
    _chk(qkv, torch.float16, "qkv")
    out = torch.empty(qkv.shape[:-1] + (128,), dtype=torch.float16, device=qkv.device)
    _lib.check(_lib.lib().wdno_linear_attn(_p(qkv), _p(out), n_img, n_pos, float(scale), _st()), "linear_attn")
    return out


# ---------------------------------------------------------------- diffusion-step algebra
class CondProgram:
    """The reference's in-place slice assignments on the state [B,F,C,H,W], in order (later overrides earlier)."""

    def __init__(self):
        self.ops = []
        self.keep = []

    def zero(self, f=(0, None), c=(0, None), y=(0, None), x=(0, None)):
        self.ops.append((f, c, y, x, None, None))
        return self

    def copy(self, src, src_dims, f=(0, None), c=(0, None), y=(0, None), x=(0, None)):
        """src: fp32 CUDA tensor; src_dims: string over 'bfcyx' naming src's dims in order, e.g. 'bfyx'."""
        _chk(src, torch.float32, "condition source")
        self.ops.append((f, c, y, x, src, src_dims))
        self.keep.append(src)
        return self

    def build(self, F, Cc, H, W):
        assert len(self.ops) <= MAX_COND_OPS, "too many condition ops"
        arr = (CondOp * max(1, len(self.ops)))()
        lim = dict(f=F, c=Cc, y=H, x=W)

        def rng(r, n):
            lo, hi = r
            lo = 0 if lo is None else (lo + n if lo < 0 else lo)
            hi = n if hi is None else (hi + n if hi < 0 else hi)
            return max(0, min(lo, n)), max(0, min(hi, n))

        for i, (f, c, y, x, src, dims) in enumerate(self.ops):
            o = arr[i]
            o.f0, o.f1 = rng(f, F)
            o.c0, o.c1 = rng(c, Cc)
            o.y0, o.y1 = rng(y, H)
            o.x0, o.x1 = rng(x, W)
            if src is not None:
                o.src = src.data_ptr()
                st = dict(zip(dims, src.stride()))
                o.sb, o.sf, o.sc, o.sy, o.sx = (st.get(k, 0) for k in "bfcyx")
                o.of, o.oc, o.oy, o.ox = o.f0, o.c0, o.y0, o.x0
                for k, (lo, hi) in zip("fcyx", ((o.f0, o.f1), (o.c0, o.c1), (o.y0, o.y1), (o.x0, o.x1))):
                    if k in dims:
                        assert src.shape[dims.index(k)] >= hi - lo, f"condition source too small along {k}"
        return arr, len(self.ops)


def _state_dims(x):
    if x.dim() == 4:
        B, Cc, H, W = x.shape
        return B, 1, Cc, H, W
    return tuple(x.shape)


def ddim_step(x, eps, noise, coef_dev, prog, cond_mode, guidance=None):
    _chk(x, torch.float32, "x")
    _chk(eps, torch.float32, "eps")
    B, F, Cc, H, W = _state_dims(x)
    arr, n = prog
    _lib.check(_lib.lib().wdno_ddim_step(_p(x), _p(eps), _p(noise), _p(guidance), _p(coef_dev), arr, n, B, F, Cc, H, W,
                                        cond_mode, _st()), "ddim_step")


def ddpm_step(x, eps, noise, coef_dev, prog, cond_mode, guidance=None):
    _chk(x, torch.float32, "x")
    _chk(eps, torch.float32, "eps")
    B, F, Cc, H, W = _state_dims(x)
    arr, n = prog
    _lib.check(_lib.lib().wdno_ddpm_step(_p(x), _p(eps), _p(noise), _p(guidance), _p(coef_dev), arr, n, B, F, Cc, H, W,
                                        cond_mode, _st()), "ddpm_step")


def apply_conditions(x, prog):
    _chk(x, torch.float32, "x")
    B, F, Cc, H, W = _state_dims(x)
    arr, n = prog
    _lib.check(_lib.lib().wdno_apply_conditions(_p(x), arr, n, B, F, Cc, H, W, _st()), "apply_conditions")


def predict_x0(x, eps, coef_dev, clip=True):
    out = torch.empty_like(x)
    _lib.check(_lib.lib().wdno_predict_x0(_p(x), _p(eps), _p(coef_dev), _p(out), x.numel(), int(clip), _st()), "predict_x0")
    return out


def q_sample(x0, noise, sqrt_ac, sqrt_1mac, t):
    _chk(x0, torch.float32, "x0")
    _chk(noise, torch.float32, "noise")
    _chk(t, torch.int64, "t")
    out = torch.empty_like(x0)
    B = x0.shape[0]
    _lib.check(_lib.lib().wdno_q_sample(_p(x0), _p(noise), _p(sqrt_ac), _p(sqrt_1mac), _p(t), _p(out), B, x0.numel() // B,
                                       _st()), "q_sample")
    return out


def mse_weighted(pred, target, w):
    """-> double [B]: sum over each sample of (pred-target)^2 * w[channel]"""
    B, F, Cc, H, W = _state_dims(pred)
    acc = torch.zeros(B, dtype=torch.float64, device=pred.device)
    wl = 0 if w is None else w.numel()
    _lib.check(_lib.lib().wdno_mse_weighted(_p(pred), _p(target), _p(w), wl, B, F, Cc, H, W, _p(acc), _st()), "mse_weighted")
    return acc


def step_begin(step_dev, time_table, coef_table, time_out, coef_out, n_steps):
    _lib.check(_lib.lib().wdno_step_begin(_p(step_dev), _p(time_table), _p(coef_table), _p(time_out), _p(coef_out),
                                         time_out.shape[0], n_steps, _st()), "step_begin")


# ---------------------------------------------------------------- sharded noise (row slice of a full-batch torch.randn)
_RANDN_ROWS_OK = {}


def _randn_rows_kernel(shape_local, full_batch, lo, device):
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    per = 1
    for d in shape_local[1:]:
        per *= int(d)
    numel_full = full_batch * per
    props = torch.cuda.get_device_properties(device)
    grid = min(props.multi_processor_count * (props.max_threads_per_multi_processor // 256), (numel_full + 255) // 256)
    out = torch.empty(tuple(shape_local), dtype=torch.float32, device=device)
    seed, off = gen.initial_seed(), gen.get_offset()
    _lib.check(_lib.lib().wdno_randn_slice(_p(out), out.numel(), lo * per, numel_full, grid, seed, off, _st()), "randn_slice")
    gen.set_offset(off + ((numel_full - 1) // (256 * grid * 4) + 1) * 4)
    return out


def _randn_rows_selfcheck(device):
    """one-time check per device that the slice kernel reproduces this torch build's normal_ (values AND generator advance),
    in both grid regimes (small tensor: grid = ceil(numel/256); large: grid = SMs * 8)"""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    state = gen.get_state()
    ok = True
    try:
        for full, per, lo, hi in ((5, 1000, 1, 4), (7, 300001, 2, 5)):
            gen.manual_seed(1234567)
            torch.randn(3, device=device)   # non-zero starting offset
            st0 = gen.get_state()
            want = torch.randn((full, per), device=device)
            off_want = gen.get_offset()
            tail_want = torch.randn(8, device=device)
            gen.set_state(st0)
            got = _randn_rows_kernel((hi - lo, per), full, lo, device)
            off_got = gen.get_offset()
            tail_got = torch.randn(8, device=device)
            ok = ok and bool(torch.equal(got, want[lo:hi])) and off_got == off_want and bool(torch.equal(tail_got, tail_want))
    finally:
        gen.set_state(state)
    return ok


def randn_rows(shape_local, full_batch, lo, device):
    """== torch.randn((full_batch,) + shape_local[1:], device=device)[lo:lo + shape_local[0]] including the generator
    advance, drawing only the requested rows (csrc/rng.cu).  Falls back to draw-and-slice if the mapping self-check fails."""
    device = torch.device(device)
    if device.type == "cuda":
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in _RANDN_ROWS_OK:
            try:
                _RANDN_ROWS_OK[key] = _randn_rows_selfcheck(device)
            except Exception:  # noqa: BLE001 - a torch build without Generator.get_offset etc.
                _RANDN_ROWS_OK[key] = False
        if _RANDN_ROWS_OK[key]:
            return _randn_rows_kernel(tuple(shape_local), full_batch, lo, device)
    return torch.randn((full_batch,) + tuple(shape_local[1:]), device=device)[lo:lo + shape_local[0]]
