"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the smoke denoiser `Unet3D_with_Conv3D`.

Functional re-statement (driven by the reference `state_dict` key names, any float dtype) of
  /root/reference/smoke/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py:357-574
and the blocks it uses (RelativePositionBias 74-112, SinusoidalPosEmb 139-151, LayerNorm 165-174,
Block/ResnetBlock 189-230, SpatialLinearAttention 232-258, EinopsToAndFrom+Attention 262-353) plus
rotary-embedding-torch's RotaryEmbedding.
Parity pin: tests/test_oracle_vs_reference.py runs it against the imported reference module in the build
container, and tests/golden/*.pt holds reference-generated vectors for the GPU box.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math

import torch
import torch.nn.functional as F


def _rel_pos_bucket(rel, num_buckets=32, max_distance=32):
    """T5-style bucket of (k_pos - q_pos); conv3d.py:86-104."""
    n = -rel
    half = num_buckets // 2
    ret = (n < 0).long() * half
    n = n.abs()
    max_exact = half // 2
    small = n < max_exact
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (half - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, half - 1))
    return ret + torch.where(small, n, large)


def rel_pos_bias(emb_weight, n, max_distance=32):
    """-> [heads, n, n]; emb_weight = time_rel_pos_bias.relative_attention_bias.weight [32, heads]."""
    pos = torch.arange(n)   # buckets on the host: float log at bucket boundaries must not depend on the device
    rel = pos[None, :] - pos[:, None]
    bucket = _rel_pos_bucket(rel, emb_weight.shape[0], max_distance).to(emb_weight.device)
    return emb_weight[bucket].permute(2, 0, 1)


def rotary(t, freqs):
    """t [..., n, d]; interleaved-pair rotation by position*freq."""
    n = t.shape[-2]
    ang = torch.arange(n, dtype=t.dtype, device=t.device)[:, None] * freqs.to(t.dtype)[None, :]
    ang = ang.repeat_interleave(2, dim=-1)
    pair = t.reshape(*t.shape[:-1], -1, 2)
    rot = torch.stack((-pair[..., 1], pair[..., 0]), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def sinusoid(time, dim):
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=time.dtype, device=time.device) * -(math.log(10000) / (half - 1)))
    e = time[:, None] * f[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def chan_layernorm(x, gamma, eps=1e-5):
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * gamma


class Unet3DOracle:
    def __init__(self, state_dict, *, heads=4, dim_head=32, groups=8, prefix="", dtype=torch.float32):
        self.sd = {k[len(prefix):]: v.detach().to(dtype) for k, v in state_dict.items() if k.startswith(prefix)}
        self.heads, self.dim_head, self.groups, self.dtype = heads, dim_head, groups, dtype
        self.n_down = len({k.split(".")[1] for k in self.sd if k.startswith("downs.")})
        self.n_up = len({k.split(".")[1] for k in self.sd if k.startswith("ups.")})
        self.dim = self.sd["time_mlp.1.weight"].shape[1]
        self.channels = self.sd["init_conv.weight"].shape[1]

    # ---- blocks
    def _block(self, x, p, scale_shift=None):
        sd = self.sd
        x = F.conv3d(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"], padding=1)
        x = F.group_norm(x, self.groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
        if scale_shift is not None:
            scale, shift = scale_shift
            x = x * (scale + 1) + shift
        return F.silu(x)

    def _resnet(self, x, p, t):
        sd = self.sd
        ss = None
        if (p + ".mlp.1.weight") in sd:
            e = F.linear(F.silu(t), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
            e = e[:, :, None, None, None]
            ss = e.chunk(2, dim=1)
        h = self._block(x, p + ".block1", ss)
        h = self._block(h, p + ".block2")
        if (p + ".res_conv.weight") in sd:
            x = F.conv3d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
        return h + x

    def _spatial_linear_attn(self, x, p):
        """Residual(PreNorm(SpatialLinearAttention)); p = '<...>.2'"""
        sd, hN = self.sd, self.heads
        b, c, f, h, w = x.shape
        xn = chan_layernorm(x, sd[p + ".fn.norm.gamma"])
        y = xn.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        qkv = F.conv2d(y, sd[p + ".fn.fn.to_qkv.weight"])
        q, k, v = [u.reshape(b * f, hN, -1, h * w) for u in qkv.chunk(3, dim=1)]
        q = q.softmax(dim=-2) * (self.dim_head ** -0.5)
        k = k.softmax(dim=-1)
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b * f, -1, h, w)
        out = F.conv2d(out, sd[p + ".fn.fn.to_out.weight"], sd[p + ".fn.fn.to_out.bias"])
        out = out.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)
        return out + x

    def _attention(self, tokens, p, pos_bias=None, use_rotary=False):
        """tokens [..., n, c] -> [..., n, c]; p = prefix of the Attention module (to_qkv/to_out)."""
        sd, hN = self.sd, self.heads
        qkv = F.linear(tokens, sd[p + ".to_qkv.weight"]).chunk(3, dim=-1)
        n = tokens.shape[-2]
        q, k, v = [u.reshape(*u.shape[:-1], hN, -1).transpose(-2, -3) for u in qkv]  # [..., h, n, d]
        q = q * (self.dim_head ** -0.5)
        if use_rotary:
            fr = sd[p + ".rotary_emb.freqs"]
            q, k = rotary(q, fr), rotary(k, fr)
        sim = q @ k.transpose(-1, -2)
        if pos_bias is not None:
            sim = sim + pos_bias
        attn = (sim - sim.amax(dim=-1, keepdim=True)).softmax(dim=-1)
        out = (attn @ v).transpose(-2, -3).reshape(*tokens.shape[:-1], -1)
        return F.linear(out, sd[p + ".to_out.weight"])

    def _temporal_attn(self, x, p, pos_bias):
        """Residual(PreNorm(EinopsToAndFrom('b c f h w','b (h w) f c', Attention)))"""
        b, c, f, h, w = x.shape
        xn = chan_layernorm(x, self.sd[p + ".fn.norm.gamma"])
        tok = xn.permute(0, 3, 4, 2, 1).reshape(b, h * w, f, c)
        out = self._attention(tok, p + ".fn.fn.fn", pos_bias, use_rotary=True)
        return out.reshape(b, h, w, f, c).permute(0, 4, 3, 1, 2) + x

    def _mid_spatial_attn(self, x, p):
        b, c, f, h, w = x.shape
        xn = chan_layernorm(x, self.sd[p + ".fn.norm.gamma"])
        tok = xn.permute(0, 2, 3, 4, 1).reshape(b, f, h * w, c)
        out = self._attention(tok, p + ".fn.fn.fn")
        return out.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3) + x

    # ---- forward
    def __call__(self, x, time, taps=None):
        """x [B, F, C, H, W], time [B] (long or float) -> [B, F, C, H, W].  `taps`: optional dict that
        receives named intermediate activations (NCDHW) for per-layer parity checks."""
        sd = self.sd
        rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
        x = x.to(self.dtype).permute(0, 2, 1, 3, 4)
        bias = rel_pos_bias(sd["time_rel_pos_bias.relative_attention_bias.weight"], x.shape[2])
        kk = sd["init_conv.weight"].shape[-1]
        x = F.conv3d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=kk // 2)
        rec("init_conv", x)
        x = self._temporal_attn(x, "init_temporal_attn", bias)
        rec("init_temporal_attn", x)
        r = x
        t = sinusoid(time.to(self.dtype), self.dim)
        t = F.linear(t, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
        t = F.gelu(t)
        t = F.linear(t, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])
        rec("time_emb", t)
        skips = []
        for i in range(self.n_down):
            p = f"downs.{i}"
            x = self._resnet(x, p + ".0", t)
            rec(p + ".0", x)
            x = self._resnet(x, p + ".1", t)
            rec(p + ".1", x)
            x = self._spatial_linear_attn(x, p + ".2")
            rec(p + ".2", x)
            x = self._temporal_attn(x, p + ".3", bias)
            rec(p + ".3", x)
            skips.append(x)
            if (p + ".4.weight") in sd:
                x = F.conv3d(x, sd[p + ".4.weight"], sd[p + ".4.bias"], stride=(1, 2, 2), padding=(0, 1, 1))
                rec(p + ".4", x)
        x = self._resnet(x, "mid_block1", t)
        rec("mid_block1", x)
        x = self._mid_spatial_attn(x, "mid_spatial_attn")
        rec("mid_spatial_attn", x)
        x = self._temporal_attn(x, "mid_temporal_attn", bias)
        rec("mid_temporal_attn", x)
        x = self._resnet(x, "mid_block2", t)
        rec("mid_block2", x)
        for i in range(self.n_up):
            p = f"ups.{i}"
            x = torch.cat((x, skips.pop()), dim=1)
            x = self._resnet(x, p + ".0", t)
            rec(p + ".0", x)
            x = self._resnet(x, p + ".1", t)
            rec(p + ".1", x)
            x = self._spatial_linear_attn(x, p + ".2")
            rec(p + ".2", x)
            x = self._temporal_attn(x, p + ".3", bias)
            rec(p + ".3", x)
            if (p + ".4.weight") in sd:
                x = F.conv_transpose3d(x, sd[p + ".4.weight"], sd[p + ".4.bias"], stride=(1, 2, 2), padding=(0, 1, 1))
                rec(p + ".4", x)
        x = torch.cat((x, r), dim=1)
        x = self._resnet(x, "final_conv.0", None)
        rec("final_conv.0", x)
        x = F.conv3d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
        return x.permute(0, 2, 1, 3, 4)
