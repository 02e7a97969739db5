"""Differentiable Burgers Unet2D forward on the engine (training step; reference forward unet.py:372-411, training loop
burgers/ddpm_burgers/train_diffusion.py:187-237).  Same machinery as train3d.py (2-D layers are depth-1 3-D layers): tap-GEMM
dgrad, wgrad, GroupNorm(1 group)/SiLU backward, attention-core backward kernels; PixelUnshuffle + 1x1 and nearest-up + 3x3 are
handled as phases of the strided kinds (training.py::ConvLayer)."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .train3d import Unet3DTrainEngine, _Unet3DFn, flat_grads
from .training import AttnGrad, ConvLayer


class Unet2DTrainEngine(Unet3DTrainEngine):
    def __init__(self, model):
        self.m = model
        self.eng = model.engine()
        e, m = self.eng, model
        self.groups = e.groups
        self.layers, self.attn = {}, {}
        dev = e.dev

        def reg(fwd, weight, bias, kind="conv", need_dgrad=True):
            self.layers[id(fwd)] = ConvLayer(fwd, weight, bias, kind, fwd.src_channels, need_dgrad)

        def reg_resnet(rp, blk):
            rp.mod = blk
            reg(rp.conv1, blk.block1.proj.weight, blk.block1.proj.bias)
            reg(rp.conv2, blk.block2.proj.weight, blk.block2.proj.bias)
            if rp.res is not None:
                reg(rp.res, blk.res_conv.weight, blk.res_conv.bias)

        def reg_lattn(ap, mod):
            a = mod.fn.fn
            ap["mod"] = mod
            self.attn[id(ap["qkv"])] = AttnGrad("linear", mod.fn.norm.g, a.to_qkv, a.to_out[0], dev, out_norm=a.to_out[1].g)

        reg(e.init_conv, m.init_conv.weight, m.init_conv.bias, need_dgrad=False)
        for lv, mods in zip(e.downs, m.downs):
            reg_resnet(lv["b1"], mods[0])
            reg_resnet(lv["b2"], mods[1])
            reg_lattn(lv["attn"], mods[2])
            if isinstance(mods[3], nn.Sequential):
                reg(lv["down"], mods[3][1].weight, mods[3][1].bias, kind="unshuffle")
            else:
                reg(lv["down"], mods[3].weight, mods[3].bias)
        reg_resnet(e.mid1, m.mid_block1)
        reg_resnet(e.mid2, m.mid_block2)
        a = m.mid_attn.fn.fn
        e.mid_attn["mod"] = m.mid_attn
        self.attn[id(e.mid_attn["qkv"])] = AttnGrad("spatial", m.mid_attn.fn.norm.g, a.to_qkv, a.to_out, dev)
        for lv, mods in zip(e.ups, m.ups):
            reg_resnet(lv["b1"], mods[0])
            reg_resnet(lv["b2"], mods[1])
            reg_lattn(lv["attn"], mods[2])
            conv = mods[3][1] if isinstance(mods[3], nn.Sequential) else mods[3]
            reg(lv["up"], conv.weight, conv.bias)
        reg_resnet(e.final_block, m.final_res_block)
        reg(e.final_conv, m.final_conv.weight, m.final_conv.bias)
        self._graphed = None
        self.use_torch_attention = False
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

    def record_params(self, rec):
        if rec[0] in ("lattn", "mattn"):
            return [p for _, p in rec[1]["mod"].named_parameters() if p.requires_grad]
        return super().record_params(rec)

    def forward(self, x, time, ss):
        """x [B,C,H,W] fp32 -> eps [B,out_dim,H,W]; records the tape (see Unet2DEngine.forward for the launch order)"""
        e = self.eng
        assert e is self.m.engine(), "parameters were re-allocated: build a new training engine"
        self.refresh()
        tape, G = [], self.groups
        B, Cc, H, W = x.shape
        x = x.contiguous().float()
        ss = ss.contiguous()
        self.d_ss = torch.zeros_like(ss)

        def resnet(rp, s0, s1):
            D_, H_, W_ = s0.shape[1:4]
            count = float(D_ * H_ * W_ * (rp.cout // G))
            st = torch.zeros((2, B, G, 2), dtype=torch.float64, device=x.device)
            y1 = rp.conv1(s0, s1, stats=st[0], groups=G)
            a1, c1 = ops.gn_finalize(st[0], rp.g1, rp.b1, ss, rp.ss_off, ss.shape[1], B, rp.cout, G, count)
            y2 = rp.conv2(y1, coef0=(a1, c1), stats=st[1], groups=G)
            a2, c2 = ops.gn_finalize(st[1], rp.g2, rp.b2, None, 0, 0, B, rp.cout, G, count)
            if rp.res is None:
                out = ops.gn_silu_add(y2, a2, c2, resid=s0)
            else:
                out = rp.res(s0, s1, resid=ops.gn_silu_add(y2, a2, c2, resid=None))
            tape.append(("resnet", rp, s0, s1, y1, a1, c1, y2, a2, c2, st, count, out))
            return out

        def conv(plan, h):
            o = plan(h)
            tape.append(("conv", plan, (h,), o))
            return o

        def lattn(ap, h):
            o = e._linear_attn(ap, h)
            tape.append(("lattn", ap, h, o))
            return o

        xin = ops.pack_bfchw_f16(x.reshape(B, 1, Cc, H, W), e.cin_pad)
        h = conv(e.init_conv, xin)
        r = h
        skips = []
        for lv in e.downs:
            h = resnet(lv["b1"], h, None)
            skips.append(h)
            h = resnet(lv["b2"], h, None)
            h = lattn(lv["attn"], h)
            skips.append(h)
            h = conv(lv["down"], h)
        h = resnet(e.mid1, h, None)
        o = e._mid_attn(e.mid_attn, h)
        tape.append(("mattn", e.mid_attn, h, o))
        h = resnet(e.mid2, o, None)
        for lv in e.ups:
            h = resnet(lv["b1"], h, skips.pop())
            h = resnet(lv["b2"], h, skips.pop())
            h = lattn(lv["attn"], h)
            h = conv(lv["up"], h)
        h = resnet(e.final_block, h, r)
        out = e.final_conv(h, out_fp32_bfchw=True)     # [B, 1, C, H, W]
        tape.append(("final", e.final_conv, (h,), out))
        self.tape = tape
        return out.reshape(B, -1, H, W)

    def backward(self, d_eps):
        B, Cc, H, W = d_eps.shape
        return super().backward(d_eps.reshape(B, 1, Cc, H, W))


def time_embedding_ss_2d(model, time):
    """SinusoidalPosEmb(dim, theta) -> time_mlp -> every ResnetBlock.mlp, concatenated in engine order (unet.py:96-111,
    304-309,150-157); plain torch ops under autograd"""
    half = model.dim // 2
    f = torch.exp(torch.arange(half, device=time.device, dtype=torch.float32) * -(math.log(model.theta) / (half - 1)))
    e = time.float()[:, None] * f[None, :]
    emb = torch.cat((e.sin(), e.cos()), dim=-1)
    emb = model.time_mlp[3](F.gelu(model.time_mlp[1](emb)))
    act = F.silu(emb)
    blocks = []
    for mods in model.downs:
        blocks += [mods[0], mods[1]]
    blocks += [model.mid_block1, model.mid_block2]
    for mods in model.ups:
        blocks += [mods[0], mods[1]]
    blocks.append(model.final_res_block)
    return torch.cat([blk.mlp[1](act) for blk in blocks], dim=1)


def unet2d_apply(model, x, time):
    """differentiable eps = model(x, time) for the Burgers Unet2D: forward and backward in libwdno_b200.so"""
    te = getattr(model, "_train_engine", None)
    if te is None or te.eng is not model.engine():
        te = Unet2DTrainEngine(model)
        model._train_engine = te
    flat_grads(model)
    te._params = [p for p in model.parameters() if p.requires_grad]
    ss = time_embedding_ss_2d(model, time)
    return _Unet3DFn.apply(x, time, ss, te, *te._params)
