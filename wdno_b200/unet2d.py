"""Burgers Unet2D on the B200 engine: same constructor / `state_dict()` keys as
/root/reference/burgers/ddpm_burgers/unet.py:263-370 and `forward(x, time)` (372-411) executed by libwdno_b200.so:
the same tap-GEMM / GroupNorm / attention kernels as the smoke Unet3D with D = 1 (2-D convolutions are depth-1 3-D
convolutions; PixelUnshuffle+1x1 and nearest-up+3x3 are operand-load modes of the tap-GEMM).
`Unet1D` (unet.py:414-549) is never constructed by any WDNO script; the name is exported for import compatibility and
raises on construction.
"""
import torch
from torch import nn

from . import ops
from ._engine_cache import EngineOwner
from .tapgemm import TapGemm


def _holder():
    return nn.Module()


def _block(dim, dim_out, groups):
    m = _holder()
    m.proj = nn.Conv2d(dim, dim_out, 3, padding=1)
    m.norm = nn.GroupNorm(groups, dim_out)
    m.act = nn.SiLU()
    return m


def _resnet(dim, dim_out, time_emb_dim, groups):
    m = _holder()
    m.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2))
    m.block1 = _block(dim, dim_out, groups)
    m.block2 = _block(dim_out, dim_out, groups)
    m.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
    return m


def _ln(dim):
    m = _holder()
    m.g = nn.Parameter(torch.ones(1, dim, 1, 1))
    return m


def _prenorm_residual(dim, fn):
    pre = _holder()
    pre.fn = fn
    pre.norm = _ln(dim)
    res = _holder()
    res.fn = pre
    return res


def _linear_attention(dim, heads=4, dim_head=32):
    m = _holder()
    hidden = heads * dim_head
    m.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
    m.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), _ln(dim))
    return m


def _attention(dim, heads=4, dim_head=32):
    m = _holder()
    hidden = heads * dim_head
    m.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
    m.to_out = nn.Conv2d(hidden, dim, 1)
    return m


class Unet2D(EngineOwner, nn.Module):
    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=2, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False,
                 random_fourier_features=False, learned_sinusoidal_dim=16, sinusoidal_pos_emb_theta=10000,
                 attn_dim_head=32, attn_heads=4):
        super().__init__()
        if self_condition or learned_sinusoidal_cond or random_fourier_features or learned_variance:
            raise NotImplementedError("WDNO constructs Unet2D with the defaults for these flags (train_ddpm_burgers.py:149-156)")
        if attn_heads != 4 or attn_dim_head != 32:
            raise NotImplementedError("attention kernels are built for heads=4, dim_head=32 (the reference defaults)")
        self.channels = channels
        self.self_condition = False
        self.random_or_learned_sinusoidal_cond = False
        self.dim, self.theta, self.groups = dim, sinusoidal_pos_emb_theta, resnet_block_groups
        time_dim = dim * 4
        self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(dim, time_dim), nn.GELU(), nn.Linear(time_dim, time_dim))
        init_dim = init_dim if init_dim is not None else dim
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        g = resnet_block_groups
        self.downs = nn.ModuleList([])
        for ind, (di, do) in enumerate(in_out):
            last = ind >= len(in_out) - 1
            # construction order == reference order, so the same seed reproduces the reference's default init
            mods = [_resnet(di, di, time_dim, g), _resnet(di, di, time_dim, g), _prenorm_residual(di, _linear_attention(di))]
            mods.append(nn.Sequential(nn.Identity(), nn.Conv2d(di * 4, do, 1)) if not last else nn.Conv2d(di, do, 3, padding=1))
            self.downs.append(nn.ModuleList(mods))
        mid = dims[-1]
        self.mid_block1 = _resnet(mid, mid, time_dim, g)
        self.mid_attn = _prenorm_residual(mid, _attention(mid))
        self.mid_block2 = _resnet(mid, mid, time_dim, g)
        self.ups = nn.ModuleList([])
        for ind, (di, do) in enumerate(reversed(in_out)):
            last = ind == len(in_out) - 1
            mods = [_resnet(do + di, do, time_dim, g), _resnet(do + di, do, time_dim, g),
                    _prenorm_residual(do, _linear_attention(do))]
            mods.append(nn.Sequential(nn.Identity(), nn.Conv2d(do, di, 3, padding=1)) if not last else nn.Conv2d(do, di, 3, padding=1))
            self.ups.append(nn.ModuleList(mods))
        self.out_dim = out_dim if out_dim is not None else channels
        self.final_res_block = _resnet(dim * 2, dim, time_dim, g)
        self.final_conv = nn.Conv2d(dim, self.out_dim, 1)
        self._engine = None

    def _make_engine(self):
        return Unet2DEngine(self)

    def forward(self, x, time, x_self_cond=None):
        """x [B, C, H, W] fp32 CUDA, time [B] -> eps [B, out_dim, H, W] fp32"""
        if not x.is_cuda:
            raise RuntimeError("wdno_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        return self.engine().forward(x, time)


class Unet1D(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("Unet1D is never constructed by WDNO (SURVEY.md section 0); the Burgers scripts use Unet2D")


class _RP:
    pass


class Unet2DEngine:
    def __init__(self, m):
        dev = next(m.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("wdno_b200: move the model to a CUDA device before calling it")
        self.dev, self.groups, self.scale = dev, m.groups, 32 ** -0.5
        self.dim, self.theta = m.dim, float(m.theta)
        f32 = self._f32
        self.cin = m.channels
        self.cin_pad = (m.channels + 15) // 16 * 16
        self._refreshers, self._plans = [], []
        self.init_conv = self._conv(m.init_conv.weight, m.init_conv.bias, src_channels=(self.cin_pad,))
        self.tw1, self.tb1 = f32(m.time_mlp[1].weight), f32(m.time_mlp[1].bias)
        self.tw2, self.tb2 = f32(m.time_mlp[3].weight), f32(m.time_mlp[3].bias)
        self._mlp_w, self._mlp_b, self._mlp_off, self.stats_slots = [], [], 0, 0
        self._mlp_w_src, self._mlp_b_src = [], []
        self.downs = []
        for b1, b2, attn, down in m.downs:
            if isinstance(down, nn.Sequential):
                dn = self._conv(down[1].weight, down[1].bias, kind="unshuffle")
            else:
                dn = self._conv(down.weight, down.bias)
            self.downs.append(dict(b1=self._resnet_plan(b1, None), b2=self._resnet_plan(b2, None),
                                   attn=self._lin_attn_plan(attn), down=dn))
        self.mid1 = self._resnet_plan(m.mid_block1, None)
        a = m.mid_attn.fn.fn
        self.mid_attn = dict(g=f32(m.mid_attn.fn.norm.g.reshape(-1)), qkv=self._conv(a.to_qkv.weight, None),
                             out=self._conv(a.to_out.weight, a.to_out.bias))
        self.mid2 = self._resnet_plan(m.mid_block2, None)
        self.ups = []
        for b1, b2, attn, up in m.ups:
            do = b1.block1.proj.weight.shape[0]
            di = b1.block1.proj.weight.shape[1] - do
            if isinstance(up, nn.Sequential):
                u = self._conv(up[1].weight, up[1].bias, up2=True)
            else:
                u = self._conv(up.weight, up.bias)
            self.ups.append(dict(b1=self._resnet_plan(b1, (do, di)), b2=self._resnet_plan(b2, (do, di)),
                                 attn=self._lin_attn_plan(attn), up=u))
        d = m.dim
        self.final_block = self._resnet_plan(m.final_res_block, (d, d))
        self.final_conv = self._conv(m.final_conv.weight, m.final_conv.bias)
        self.mlp_w = torch.cat(self._mlp_w, 0).contiguous()
        self.mlp_b = torch.cat(self._mlp_b, 0).contiguous()
        self.launches = 0

    def _f32(self, t):
        return t.detach().to(self.dev, torch.float32).contiguous()

    def _conv(self, weight, bias=None, **kw):
        """TapGemm of a parameter pair, registered for in-place refresh"""
        plan = TapGemm(weight, bias, device=self.dev, **kw)
        self._plans.append(plan)
        self._refreshers.append(lambda: plan.refresh(weight, bias))
        return plan

    def refresh(self):
        """re-pack every weight-derived buffer from the live parameters in place (see Unet3DEngine.refresh)"""
        from ._engine_cache import GraphedRefresh
        if getattr(self, "_graphed_refresh", None) is None:
            self._graphed_refresh = GraphedRefresh(self._refresh_eager, lambda: sum(len(p._packed) for p in self._plans))
        self._graphed_refresh()

    def _refresh_eager(self):
        for fn in self._refreshers:
            fn()
        self.mlp_w.copy_(torch.cat([w.detach().float() for w in self._mlp_w_src], 0))
        self.mlp_b.copy_(torch.cat([b.detach().float() for b in self._mlp_b_src], 0))

    def _resnet_plan(self, blk, src_channels):
        p = _RP()
        p.cout = blk.block1.proj.weight.shape[0]
        p.conv1 = self._conv(blk.block1.proj.weight, blk.block1.proj.bias, src_channels=src_channels)
        p.conv2 = self._conv(blk.block2.proj.weight, blk.block2.proj.bias)
        p.g1, p.b1 = self._f32(blk.block1.norm.weight), self._f32(blk.block1.norm.bias)
        p.g2, p.b2 = self._f32(blk.block2.norm.weight), self._f32(blk.block2.norm.bias)
        p.res = None
        if not isinstance(blk.res_conv, nn.Identity):
            p.res = self._conv(blk.res_conv.weight, blk.res_conv.bias, src_channels=src_channels)
        p.ss_off = self._mlp_off
        self._mlp_w.append(self._f32(blk.mlp[1].weight))
        self._mlp_b.append(self._f32(blk.mlp[1].bias))
        self._mlp_w_src.append(blk.mlp[1].weight)
        self._mlp_b_src.append(blk.mlp[1].bias)
        self._mlp_off += 2 * p.cout
        p.stat1, p.stat2 = self.stats_slots, self.stats_slots + 1
        self.stats_slots += 2
        return p

    def _lin_attn_plan(self, res):
        a = res.fn.fn
        return dict(g=self._f32(res.fn.norm.g.reshape(-1)), qkv=self._conv(a.to_qkv.weight, None),
                    out=self._conv(a.to_out[0].weight, a.to_out[0].bias), g_out=self._f32(a.to_out[1].g.reshape(-1)))

    def _resnet(self, p, src0, src1, ss, stats):
        B, D, H, W, _ = src0.shape
        G = self.groups
        count = float(D * H * W * (p.cout // G))
        st1, st2 = stats[p.stat1], stats[p.stat2]
        y1 = p.conv1(src0, src1, stats=st1, groups=G)
        a1, c1 = ops.gn_finalize(st1, p.g1, p.b1, ss, p.ss_off, ss.shape[1] if self._ss_stride is None else self._ss_stride,
                                 B, p.cout, G, count)
        y2 = p.conv2(y1, coef0=(a1, c1), stats=st2, groups=G)
        a2, c2 = ops.gn_finalize(st2, p.g2, p.b2, None, 0, 0, B, p.cout, G, count)
        self.launches += 5
        if p.res is None:
            return ops.gn_silu_add(y2, a2, c2, resid=src0)
        self.launches += 1
        return p.res(src0, src1, resid=ops.gn_silu_add(y2, a2, c2, resid=None))

    def _linear_attn(self, ap, x):
        B, D, H, W, C = x.shape
        qkv = ap["qkv"](ops.chan_layernorm(x, ap["g"]))
        o = ops.linear_attn(qkv, B * D, H * W, self.scale)
        y = ap["out"](o)
        self.launches += 5
        return ops.chan_layernorm(y, ap["g_out"], resid=x)

    def _mid_attn(self, ap, x):
        B, D, H, W, C = x.shape
        qkv = ap["qkv"](ops.chan_layernorm(x, ap["g"]))
        o = ops.softmax_attn(qkv, B * D, H * W, 1, H * W, 0, 1, self.scale)
        self.launches += 4
        return ap["out"](o, resid=x)

    def forward(self, x, time, taps=None):
        rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
        self.launches = 0
        x = x.contiguous().float()
        B, C, H, W = x.shape
        tf = time.to(device=x.device, dtype=torch.float32).contiguous()
        self._ss_stride = None  # see Unet3DEngine.forward: batch-uniform time -> one embedded row, zero row stride
        if getattr(self, "time_uniform", False):
            tf = tf[:1]
            self._ss_stride = 0
        emb, emb_silu = ops.time_mlp(tf, self.tw1, self.tb1, self.tw2, self.tb2, theta=self.theta)
        ss = ops.small_linear(emb_silu, self.mlp_w, self.mlp_b)
        stats = torch.zeros((self.stats_slots, B, self.groups, 2), dtype=torch.float64, device=x.device)
        xin = ops.pack_bfchw_f16(x.reshape(B, 1, C, H, W), self.cin_pad)
        h = self.init_conv(xin)
        self.launches += 5
        rec("init_conv", h)
        r = h
        skips = []
        for i, lv in enumerate(self.downs):
            h = self._resnet(lv["b1"], h, None, ss, stats)
            rec(f"downs.{i}.0", h)
            skips.append(h)
            h = self._resnet(lv["b2"], h, None, ss, stats)
            rec(f"downs.{i}.1", h)
            h = self._linear_attn(lv["attn"], h)
            rec(f"downs.{i}.2", h)
            skips.append(h)
            h = lv["down"](h)
            self.launches += 1
            rec(f"downs.{i}.3", h)
        h = self._resnet(self.mid1, h, None, ss, stats)
        rec("mid_block1", h)
        h = self._mid_attn(self.mid_attn, h)
        rec("mid_attn", h)
        h = self._resnet(self.mid2, h, None, ss, stats)
        rec("mid_block2", h)
        for i, lv in enumerate(self.ups):
            h = self._resnet(lv["b1"], h, skips.pop(), ss, stats)
            rec(f"ups.{i}.0", h)
            h = self._resnet(lv["b2"], h, skips.pop(), ss, stats)
            rec(f"ups.{i}.1", h)
            h = self._linear_attn(lv["attn"], h)
            rec(f"ups.{i}.2", h)
            h = lv["up"](h)
            self.launches += 1
            rec(f"ups.{i}.3", h)
        h = self._resnet(self.final_block, h, r, ss, stats)
        rec("final_res_block", h)
        out = self.final_conv(h, out_fp32_bfchw=True)  # [B, 1, C, H, W]
        self.launches += 1
        return out.reshape(B, -1, H, W)
