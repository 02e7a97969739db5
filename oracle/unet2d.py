"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the Burgers denoiser `Unet2D`.

Functional re-statement (driven by the reference `state_dict` key names) of
  /root/reference/burgers/ddpm_burgers/unet.py:263-411 and its blocks (LayerNorm 55-65, SinusoidalPosEmb 82-108,
  Block/ResnetBlock 129-181, LinearAttention 183-223, Attention 225-259, Downsample2d/Upsample2d 35-45).
Pinned against the imported reference module by tests/test_oracle_vs_reference.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import math

import torch
import torch.nn.functional as F


def _ln(x, g, eps=1e-5):
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


class Unet2DOracle:
    def __init__(self, state_dict, *, groups=1, heads=4, dim_head=32, theta=10000, prefix="", dtype=torch.float32):
        self.sd = {k[len(prefix):]: v.detach().to(dtype) for k, v in state_dict.items() if k.startswith(prefix)}
        self.groups, self.heads, self.dim_head, self.theta, self.dtype = groups, heads, dim_head, theta, dtype
        self.n_down = len({k.split(".")[1] for k in self.sd if k.startswith("downs.")})
        self.n_up = len({k.split(".")[1] for k in self.sd if k.startswith("ups.")})
        self.dim = self.sd["time_mlp.1.weight"].shape[1]

    def _block(self, x, p, ss=None):
        sd = self.sd
        x = F.conv2d(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"], padding=1)
        x = F.group_norm(x, self.groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
        if ss is not None:
            x = x * (ss[0] + 1) + ss[1]
        return F.silu(x)

    def _resnet(self, x, p, t):
        sd = self.sd
        e = F.linear(F.silu(t), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])[:, :, None, None]
        h = self._block(x, p + ".block1", e.chunk(2, dim=1))
        h = self._block(h, p + ".block2")
        if (p + ".res_conv.weight") in sd:
            x = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
        return h + x

    def _linear_attn(self, x, p):
        """Residual(PreNorm(LinearAttention)) with the trailing LayerNorm of to_out"""
        sd, hN = self.sd, self.heads
        b, c, h, w = x.shape
        xn = _ln(x, sd[p + ".fn.norm.g"])
        q, k, v = [u.reshape(b, hN, -1, h * w) for u in F.conv2d(xn, sd[p + ".fn.fn.to_qkv.weight"]).chunk(3, dim=1)]
        q = q.softmax(dim=-2) * (self.dim_head ** -0.5)
        k = k.softmax(dim=-1)
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, -1, h, w)
        out = F.conv2d(out, sd[p + ".fn.fn.to_out.0.weight"], sd[p + ".fn.fn.to_out.0.bias"])
        return _ln(out, sd[p + ".fn.fn.to_out.1.g"]) + x

    def _attn(self, x, p):
        sd, hN = self.sd, self.heads
        b, c, h, w = x.shape
        xn = _ln(x, sd[p + ".fn.norm.g"])
        q, k, v = [u.reshape(b, hN, -1, h * w) for u in F.conv2d(xn, sd[p + ".fn.fn.to_qkv.weight"]).chunk(3, dim=1)]
        sim = torch.einsum("bhdi,bhdj->bhij", q * (self.dim_head ** -0.5), k)
        out = torch.einsum("bhij,bhdj->bhid", sim.softmax(dim=-1), v)
        out = out.permute(0, 1, 3, 2).reshape(b, -1, h, w)
        return F.conv2d(out, sd[p + ".fn.fn.to_out.weight"], sd[p + ".fn.fn.to_out.bias"]) + x

    def _down(self, x, p):
        sd = self.sd
        if (p + ".1.weight") in sd:  # Rearrange + 1x1 conv
            b, c, h, w = x.shape
            x = x.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, h // 2, w // 2)
            return F.conv2d(x, sd[p + ".1.weight"], sd[p + ".1.bias"])
        return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)

    def _up(self, x, p):
        sd = self.sd
        if (p + ".1.weight") in sd:  # nearest x2 + 3x3 conv
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            return F.conv2d(x, sd[p + ".1.weight"], sd[p + ".1.bias"], padding=1)
        return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)

    def __call__(self, x, time, taps=None):
        sd = self.sd
        rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
        x = x.to(self.dtype)
        x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
        rec("init_conv", x)
        r = x
        half = self.dim // 2
        f = torch.exp(torch.arange(half, dtype=self.dtype, device=x.device) * -(math.log(self.theta) / (half - 1)))
        e = time.to(self.dtype)[:, None] * f[None, :]
        t = torch.cat((e.sin(), e.cos()), dim=-1)
        t = F.linear(F.gelu(F.linear(t, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])), sd["time_mlp.3.weight"],
                     sd["time_mlp.3.bias"])
        skips = []
        for i in range(self.n_down):
            p = f"downs.{i}"
            x = self._resnet(x, p + ".0", t)
            rec(p + ".0", x)
            skips.append(x)
            x = self._resnet(x, p + ".1", t)
            rec(p + ".1", x)
            x = self._linear_attn(x, p + ".2")
            rec(p + ".2", x)
            skips.append(x)
            x = self._down(x, p + ".3")
            rec(p + ".3", x)
        x = self._resnet(x, "mid_block1", t)
        rec("mid_block1", x)
        x = self._attn(x, "mid_attn")
        rec("mid_attn", x)
        x = self._resnet(x, "mid_block2", t)
        rec("mid_block2", x)
        for i in range(self.n_up):
            p = f"ups.{i}"
            x = self._resnet(torch.cat((x, skips.pop()), dim=1), p + ".0", t)
            rec(p + ".0", x)
            x = self._resnet(torch.cat((x, skips.pop()), dim=1), p + ".1", t)
            rec(p + ".1", x)
            x = self._linear_attn(x, p + ".2")
            rec(p + ".2", x)
            x = self._up(x, p + ".3")
            rec(p + ".3", x)
        x = self._resnet(torch.cat((x, r), dim=1), "final_res_block", t)
        rec("final_res_block", x)
        return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])
