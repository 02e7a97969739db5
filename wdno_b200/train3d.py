"""Differentiable Unet3D_with_Conv3D forward on the engine (training step; reference forward conv3d.py:487-574, backward
by autograd in Trainer.train, diffusion_2d.py:1277-1284).  See wdno_b200/training.py for what runs where.

`unet3d_apply(model, x, time)` returns eps with a grad_fn; its backward fills `p.grad` of every parameter of `model`
(views into ONE flat fp32 buffer, `model._flat_grad`)."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .training import AttnGrad, ConvLayer, add_f16, chan_layernorm_bwd, gn_bwd, pack_grad_f16

HEADS, DH = 4, 32


# ---------------------------------------------------------------------------------------------- flat gradient buffer
def flat_grads(model):
    """p.grad of every trainable parameter = a view into one flat fp32 buffer (created once, reused)."""
    params = [p for p in model.parameters() if p.requires_grad]
    buf = getattr(model, "_flat_grad", None)
    n = sum(p.numel() for p in params)
    if buf is None or buf.numel() != n or buf.device != params[0].device:
        buf = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        model._flat_grad = buf
    off = 0
    for p in params:
        v = buf[off:off + p.numel()].view_as(p)
        if p.grad is None:
            v.zero_()            # optimizer.zero_grad(set_to_none=True) dropped the view: the slice starts from zero again
            p.grad = v
        elif p.grad.data_ptr() != v.data_ptr():
            v.copy_(p.grad)      # a gradient tensor from elsewhere: keep its value, move it into the flat buffer
            p.grad = v
        off += p.numel()
    return buf


# ---------------------------------------------------------------------------------------------- attention blocks in torch (backward only)
def _chan_ln(x, gamma, eps=1e-5):
    mean = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * gamma.reshape(-1)


def _rotary(t, freqs):
    n = t.shape[-2]
    ang = torch.arange(n, dtype=t.dtype, device=t.device)[:, None] * freqs.to(t.dtype)[None, :]
    ang = ang.repeat_interleave(2, dim=-1)
    pair = t.reshape(*t.shape[:-1], -1, 2)
    rot = torch.stack((-pair[..., 1], pair[..., 0]), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def _rel_bucket(n, num_buckets=32, max_distance=32):
    pos = torch.arange(n)
    k = -(pos[None, :] - pos[:, None])
    half = num_buckets // 2
    ret = (k < 0).long() * half
    k = k.abs()
    max_exact = half // 2
    large = max_exact + (torch.log(k.float() / max_exact) / math.log(max_distance / max_exact) * (half - max_exact)).long()
    large = torch.min(large, torch.full_like(large, half - 1))
    return ret + torch.where(k < max_exact, k, large)


def _softmax_attention(tok, wqkv, wout, bias=None, freqs=None):
    """tok [N, n, C] -> [N, n, C]  (Attention.forward, conv3d.py:294-353, all-False focus mask)"""
    N, n, _ = tok.shape
    q, k, v = [u.reshape(N, n, HEADS, DH).transpose(1, 2) for u in (tok @ wqkv.t()).chunk(3, dim=-1)]
    q = q * (DH ** -0.5)
    if freqs is not None:
        q, k = _rotary(q, freqs), _rotary(k, freqs)
    sim = q @ k.transpose(-1, -2)
    if bias is not None:
        sim = sim + bias
    out = (sim.softmax(dim=-1) @ v).transpose(1, 2).reshape(N, n, HEADS * DH)
    return out @ wout.t()


def temporal_block_torch(x, gamma, wqkv, wout, rel_emb, freqs):
    """x [B,D,H,W,C] fp32 -> Residual(PreNorm(EinopsToAndFrom('b c f h w','b (h w) f c', Attention)))"""
    B, D, H, W, Cc = x.shape
    tok = _chan_ln(x, gamma).permute(0, 2, 3, 1, 4).reshape(B * H * W, D, Cc)
    bias = rel_emb[_rel_bucket(D).to(rel_emb.device)].permute(2, 0, 1)
    out = _softmax_attention(tok, wqkv, wout, bias, freqs)
    return out.reshape(B, H, W, D, Cc).permute(0, 3, 1, 2, 4) + x


def mid_spatial_block_torch(x, gamma, wqkv, wout):
    B, D, H, W, Cc = x.shape
    tok = _chan_ln(x, gamma).reshape(B * D, H * W, Cc)
    return _softmax_attention(tok, wqkv, wout).reshape(B, D, H, W, Cc) + x


def linattn_block_torch(x, gamma, wqkv, wout, bout):
    """Residual(PreNorm(SpatialLinearAttention)) (conv3d.py:232-258) on [B,D,H,W,C]"""
    B, D, H, W, Cc = x.shape
    tok = _chan_ln(x, gamma).reshape(B * D, H * W, Cc)
    qkv = tok @ wqkv.reshape(wqkv.shape[0], -1).t()                       # [I, n, 384]
    q, k, v = [u.reshape(B * D, H * W, HEADS, DH).permute(0, 2, 3, 1) for u in qkv.chunk(3, dim=-1)]   # [I, h, d, n]
    q = q.softmax(dim=-2) * (DH ** -0.5)
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B * D, HEADS * DH, H * W).transpose(1, 2)   # [I, n, 128]
    out = out @ wout.reshape(wout.shape[0], -1).t() + bout
    return out.reshape(B, D, H, W, Cc) + x


def _torch_block_backward(fn, x16, params, dy16, scale):
    """gradient of a torch-evaluated block: -> dx (fp16, scaled domain); parameter gradients += (unscaled)"""
    with torch.enable_grad():
        xin = x16.float().requires_grad_(True)
        leaves = [p.detach().requires_grad_(p.requires_grad) for p in params]
        y = fn(xin, *leaves)
        wanted = [xin] + [l for l in leaves if l.requires_grad]
        grads = torch.autograd.grad(y, wanted, grad_outputs=dy16.float())
    gi = iter(grads[1:])
    for p, l in zip(params, leaves):
        if l.requires_grad:
            p.grad.add_(next(gi).reshape(p.shape), alpha=scale)
    return grads[0].clamp_(-65504.0, 65504.0).to(torch.float16)


# ---------------------------------------------------------------------------------------------- the training engine
class Unet3DTrainEngine:
    """wraps the inference engine of a Unet3D_with_Conv3D: same forward kernels, plus dgrad / wgrad plans per layer"""

    def __init__(self, model):
        self.m = model
        self.eng = model.engine()
        e = self.eng
        self.groups = e.groups
        self.layers = {}     # id(fwd plan) -> ConvLayer

        def reg(fwd, weight, bias, kind="conv", src_channels=None, need_dgrad=True):
            sc = src_channels if src_channels is not None else fwd.src_channels
            self.layers[id(fwd)] = ConvLayer(fwd, weight, bias, kind, sc, need_dgrad)

        m = model
        reg(e.init_conv, m.init_conv.weight, m.init_conv.bias, need_dgrad=False)
        for rp in self._resnet_plans():
            blk = rp.mod
            reg(rp.conv1, blk.block1.proj.weight, blk.block1.proj.bias)
            reg(rp.conv2, blk.block2.proj.weight, blk.block2.proj.bias)
            if rp.res is not None:
                reg(rp.res, blk.res_conv.weight, blk.res_conv.bias)
        for lv, mods in zip(e.downs, m.downs):
            if lv["down"] is not None:
                reg(lv["down"], mods[4].weight, mods[4].bias, kind="down144")
        for lv, mods in zip(e.ups, m.ups):
            if lv["up"] is not None:
                reg(lv["up"], mods[4].weight, mods[4].bias, kind="up144")
        reg(e.final_conv, m.final_conv[1].weight, m.final_conv[1].bias)
        # attention blocks: backward plans keyed by the forward plan object
        rel = m.time_rel_pos_bias.relative_attention_bias.weight
        self.attn = {}
        dev = e.dev

        def reg_attn(plan, mod, kind):
            a = mod.fn.fn if kind == "linear" else mod.fn.fn.fn
            key = plan["qkv"] if isinstance(plan, dict) else plan
            self.attn[id(key)] = AttnGrad(kind, mod.fn.norm.gamma, a.to_qkv, a.to_out, dev, rel_emb=rel)

        reg_attn(e.init_tattn, m.init_temporal_attn, "temporal")
        for lv, mods in list(zip(e.downs, m.downs)) + list(zip(e.ups, m.ups)):
            reg_attn(lv["sattn"], mods[2], "linear")
            reg_attn(lv["tattn"], mods[3], "temporal")
        reg_attn(e.mid_sattn, m.mid_spatial_attn, "spatial")
        reg_attn(e.mid_tattn, m.mid_temporal_attn, "temporal")
        self._graphed = None
        self.use_torch_attention = False   # True: differentiate the attention blocks with fp32 torch ops (cross-check path)
        self._versions = None

    def _resnet_plans(self):
        e = self.eng
        out = []
        for lv in e.downs:
            out += [lv["b1"], lv["b2"]]
        out += [e.mid1, e.mid2]
        for lv in e.ups:
            out += [lv["b1"], lv["b2"]]
        out.append(e.final_block)
        return out

    def refresh(self):
        """dgrad tiles follow the live parameters (the forward tiles are refreshed by model.engine())"""
        v = sum(p._version for p in self.m.parameters())
        if v != self._versions:
            if self._graphed is None:
                from ._engine_cache import GraphedRefresh
                self._graphed = GraphedRefresh(self._refresh_eager, self._refresh_signature)
            self._graphed()
            self._versions = v

    def _refresh_signature(self):
        n = sum(len(p._packed) for layer in self.layers.values() for p in layer.dgrad)
        for ag in self.attn.values():
            n += sum(len(p._packed) + len(l.fwd._packed) for l in ag.own for p in l.dgrad)
        return n

    def _refresh_eager(self):
        for layer in self.layers.values():
            layer.refresh(fwd=False)
        for ag in self.attn.values():
            ag.refresh()

    # ------------------------------------------------------------------ forward (records the tape)
    def forward(self, x, time, ss):
        """x [B,F,C,H,W] fp32, time [B], ss [B, sum 2*cout] fp32 (the ResnetBlock time-MLP outputs, computed by the caller in
        torch so that their parameters receive gradients through autograd) -> eps fp32 [B,F,C,H,W]; self.tape = backward ops"""
        e = self.eng
        assert e is self.m.engine(), "parameters were re-allocated: build a new training engine"
        self.refresh()
        tape = []
        G = self.groups
        B = x.shape[0]
        x = x.contiguous().float()
        ss = ss.contiguous()
        self.d_ss = torch.zeros_like(ss)

        def conv(plan, s0, s1=None, **kw):
            return plan(s0, s1, **kw)

        def resnet(rp, s0, s1):
            D_, H_, W_ = s0.shape[1:4]
            count = float(D_ * H_ * W_ * (rp.cout // G))
            st = torch.zeros((2, B, G, 2), dtype=torch.float64, device=x.device)
            y1 = conv(rp.conv1, s0, s1, stats=st[0], groups=G)
            has_ss = rp.ss_off is not None
            a1, c1 = ops.gn_finalize(st[0], rp.g1, rp.b1, ss if has_ss else None, rp.ss_off or 0, ss.shape[1] if has_ss else 0,
                                     B, rp.cout, G, count)
            y2 = conv(rp.conv2, y1, coef0=(a1, c1), stats=st[1], groups=G)
            a2, c2 = ops.gn_finalize(st[1], rp.g2, rp.b2, None, 0, 0, B, rp.cout, G, count)
            if rp.res is None:
                out = ops.gn_silu_add(y2, a2, c2, resid=s0)
            else:
                out = conv(rp.res, s0, s1, resid=ops.gn_silu_add(y2, a2, c2, resid=None))
            tape.append(("resnet", rp, s0, s1, y1, a1, c1, y2, a2, c2, st, count, out))
            return out

        def tattn(blk, h):
            bias, rot = e._rel_tables(h.shape[1])
            out = blk(h, bias=bias, rot=rot)
            tape.append(("tattn", blk, h, out))
            return out

        def lattn(blk, h):
            out = blk(h)
            tape.append(("lattn", blk, h, out))
            return out

        emb_unused = None  # the time embedding itself only feeds `ss`
        xin = ops.pack_bfchw_f16(x, e.cin_pad)
        h = conv(e.init_conv, xin)
        tape.append(("conv", e.init_conv, (xin,), h))
        h = tattn(e.init_tattn, h)
        r = h
        skips = []
        for lv in e.downs:
            h = resnet(lv["b1"], h, None)
            h = resnet(lv["b2"], h, None)
            h = lattn(lv["sattn"], h)
            h = tattn(lv["tattn"], h)
            skips.append(h)
            if lv["down"] is not None:
                o = conv(lv["down"], h)
                tape.append(("conv", lv["down"], (h,), o))
                h = o
        h = resnet(e.mid1, h, None)
        o = e._mid_spatial_attn(e.mid_sattn, h)
        tape.append(("mattn", e.mid_sattn, h, o))
        h = o
        h = tattn(e.mid_tattn, h)
        h = resnet(e.mid2, h, None)
        for lv in e.ups:
            h = resnet(lv["b1"], h, skips.pop())
            h = resnet(lv["b2"], h, None)
            h = lattn(lv["sattn"], h)
            h = tattn(lv["tattn"], h)
            if lv["up"] is not None:
                o = conv(lv["up"], h)
                tape.append(("conv", lv["up"], (h,), o))
                h = o
        h = resnet(e.final_block, h, r)
        out = e.final_conv(h, out_fp32_bfchw=True)
        tape.append(("final", e.final_conv, (h,), out))
        self.tape = tape
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, d_eps):
        """d_eps fp32 [B,F,C,H,W] -> d_ss (fp32, unscaled); parameter gradients += into p.grad"""
        tape, G = self.tape, self.groups
        flat_grads(self.m)
        amax = float(d_eps.abs().max())
        if not math.isfinite(amax) or amax == 0.0:
            amax = 1.0
        S = 2.0 ** round(math.log2(256.0 / amax))      # activation gradients live around 2^8 in fp16
        inv = 1.0 / S
        grads = {}

        def acc(t, g):
            k = id(t)
            grads[k] = g if k not in grads else add_f16(grads[k], g)

        rotary = self.m.init_temporal_attn.fn.fn.fn.rotary_emb.freqs if hasattr(self.m, "init_temporal_attn") else None
        rel_emb = self.m.time_rel_pos_bias.relative_attention_bias.weight if hasattr(self.m, "time_rel_pos_bias") else None
        prof = self.profile
        for rec in reversed(tape):
            kind = rec[0]
            if prof is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            if kind == "final":
                _, plan, (h,), out = rec
                layer = self.layers[id(plan)]
                dy = pack_grad_f16(d_eps.contiguous(), layer._dy_channels(), S)
                layer.backward_weight((h,), dy, inv)
                acc(h, layer.backward_input(dy, 0))
            elif kind == "conv":
                _, plan, srcs, out = rec
                layer = self.layers[id(plan)]
                dy = grads.pop(id(out))
                if layer.kind == "up144":
                    layer.backward_weight((dy,), srcs[0], inv)
                else:
                    layer.backward_weight(srcs, dy, inv)
                if layer.dgrad:
                    acc(srcs[0], layer.backward_input(dy, 0))
            elif kind == "resnet":
                _, rp, s0, s1, y1, a1, c1, y2, a2, c2, st, count, out = rec
                blk = rp.mod
                dout = grads.pop(id(out))
                l1, l2 = self.layers[id(rp.conv1)], self.layers[id(rp.conv2)]
                srcs = (s0,) if s1 is None else (s0, s1)
                dy2 = gn_bwd(dout, y2, a2, c2, st[1], rp.g2, rp.b2, blk.block2.norm.weight.grad, blk.block2.norm.bias.grad,
                             G, count, inv)
                h1 = ops.gn_silu_add(y1, a1, c1, resid=None)
                l2.backward_weight((h1,), dy2, inv)
                dh1 = l2.backward_input(dy2, 0)
                del h1
                has_ss = rp.ss_off is not None
                dy1 = gn_bwd(dh1, y1, a1, c1, st[0], rp.g1, rp.b1, blk.block1.norm.weight.grad, blk.block1.norm.bias.grad,
                             G, count, inv,
                             ss=self._ss_slice(rp) if has_ss else None, ss_stride=self._ss.shape[1] if has_ss else 0,
                             d_ss=self._dss_slice(rp) if has_ss else None, dss_stride=self.d_ss.shape[1] if has_ss else 0)
                l1.backward_weight(srcs, dy1, inv)
                if rp.res is None:
                    acc(s0, l1.backward_input(dy1, 0, resid=dout))
                else:
                    lr = self.layers[id(rp.res)]
                    lr.backward_weight(srcs, dout, inv)
                    for i, s in enumerate(srcs):
                        acc(s, l1.backward_input(dy1, i, resid=lr.backward_input(dout, i)))
            elif kind in ("tattn", "lattn", "mattn") and not self.use_torch_attention:
                _, plan, h, out = rec
                dy = grads.pop(id(out))
                tables = self.eng._rel_tables(h.shape[1]) if kind == "tattn" else None
                acc(h, self.attn[id(plan if not isinstance(plan, dict) else plan["qkv"])].backward(h, dy, inv, tables=tables))
            elif kind == "tattn":
                _, blk, h, out = rec
                mod = blk.mod
                attn = mod.fn.fn.fn
                dy = grads.pop(id(out))
                acc(h, _torch_block_backward(
                    lambda xx, g, wq, wo, re: temporal_block_torch(xx, g, wq, wo, re, rotary.detach()),
                    h, [mod.fn.norm.gamma, attn.to_qkv.weight, attn.to_out.weight, rel_emb], dy, inv))
            elif kind == "lattn":
                _, blk, h, out = rec
                mod = blk.mod
                attn = mod.fn.fn
                dy = grads.pop(id(out))
                acc(h, _torch_block_backward(linattn_block_torch, h, [mod.fn.norm.gamma, attn.to_qkv.weight, attn.to_out.weight,
                                                                    attn.to_out.bias], dy, inv))
            elif kind == "mattn":
                _, ap, h, out = rec
                mod = ap["mod"]
                attn = mod.fn.fn.fn
                dy = grads.pop(id(out))
                acc(h, _torch_block_backward(mid_spatial_block_torch, h, [mod.fn.norm.gamma, attn.to_qkv.weight, attn.to_out.weight],
                                             dy, inv))
            else:
                raise AssertionError(kind)
            if prof is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                ch = rec[2].shape[-1] if kind in ("tattn", "lattn", "mattn") else (rec[1].cout if kind == "resnet" else 0)
                prof.append((kind, ch, tuple(rec[2].shape[1:4]) if kind != "conv" and kind != "final" else (), ev0, ev1))
            if self.on_record_done is not None:
                self.on_record_done(rec)
        self.tape = None
        return self.d_ss

    profile = None          # set to a list to collect (kind, channels, grid, start event, end event) per tape record
    on_record_done = None   # optional callable(record): the trainer launches bucket all-reduces from here

    def record_params(self, rec):
        """parameters whose gradient is FINAL once `rec` has been processed by backward() (the relative-position embedding
        is shared by every temporal block and the time-embedding MLPs get their gradients from autograd afterwards: both late)"""
        kind = rec[0]
        if kind in ("final", "conv"):
            layer = self.layers[id(rec[1])]
            return [p for p in (layer.weight, layer.bias) if p is not None]
        if kind == "resnet":
            blk = rec[1].mod
            mods = [blk.block1.proj, blk.block1.norm, blk.block2.proj, blk.block2.norm]
            if not isinstance(blk.res_conv, nn.Identity):
                mods.append(blk.res_conv)
            return [p for m_ in mods for p in m_.parameters()]
        mod = rec[1]["mod"] if kind == "mattn" else rec[1].mod
        return [p for n, p in mod.named_parameters() if p.requires_grad]

    # slices of the (scale | shift) rows of one ResnetBlock inside ss / d_ss
    def _ss_slice(self, rp):
        return self._ss[:, rp.ss_off:]

    def _dss_slice(self, rp):
        return self.d_ss[:, rp.ss_off:]


class _Unet3DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, time, ss, te, *params):
        ctx.te = te
        te._ss = ss.detach().contiguous()
        with torch.no_grad():
            out = te.forward(x, time, te._ss)
        return out

    @staticmethod
    def backward(ctx, d_eps):
        te = ctx.te
        with torch.no_grad():
            d_ss = te.backward(d_eps)
        # parameter gradients were accumulated into p.grad directly (flat buffer); autograd gets None for them
        return (None, None, d_ss, None) + (None,) * (len(te._params))


def time_embedding_ss(model, time):
    """SinusoidalPosEmb -> time_mlp -> every ResnetBlock.mlp (SiLU, Linear), concatenated in engine order [B, sum 2*cout];
    plain torch ops under autograd (B x 256 values; conv3d.py:139-151,405-410,207-214)"""
    dim = model.dim
    half = dim // 2
    f = torch.exp(torch.arange(half, device=time.device, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    e = time.float()[:, None] * f[None, :]
    emb = torch.cat((e.sin(), e.cos()), dim=-1)
    emb = model.time_mlp[3](F.gelu(model.time_mlp[1](emb)))
    act = F.silu(emb)
    outs = []
    blocks = []
    for mods in model.downs:
        blocks += [mods[0], mods[1]]
    blocks += [model.mid_block1, model.mid_block2]
    for mods in model.ups:
        blocks += [mods[0], mods[1]]
    for blk in blocks:
        outs.append(blk.mlp[1](act))
    return torch.cat(outs, dim=1)


def unet3d_apply(model, x, time):
    """differentiable eps = model(x, time): forward and backward in libwdno_b200.so (see module docstring)"""
    te = getattr(model, "_train_engine", None)
    if te is None or te.eng is not model.engine():
        te = Unet3DTrainEngine(model)
        model._train_engine = te
    flat_grads(model)
    te._params = [p for p in model.parameters() if p.requires_grad]
    ss = time_embedding_ss(model, time)
    return _Unet3DFn.apply(x, time, ss, te, *te._params)
