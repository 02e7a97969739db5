"""GPU parity of the fused attention blocks (csrc/attn_fused.cu) against the fp32 oracle blocks (oracle/unet3d.py,
which is pinned bit-exactly to the reference modules in tests/test_oracle_vs_reference.py).
Tolerance: operands are fp16 (x, weights, softmax probabilities), accumulation fp32 -> relative L2 of the block's
residual branch (y - x) below 6e-3 (six chained fp16-operand products; measured 2-3e-3), and of y below 1e-3."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("C,B,D,H,W", [(64, 2, 3, 40, 40), (128, 1, 5, 20, 20), (256, 2, 2, 10, 10), (64, 1, 2, 9, 7)])
def test_linattn_block_vs_oracle(C, B, D, H, W):
    from oracle.unet3d import Unet3DOracle
    from wdno_b200.attn_fused import LinAttnBlock
    torch.manual_seed(C + H)
    sd = {
        "time_mlp.1.weight": torch.zeros(4, 1), "init_conv.weight": torch.zeros(1, 1),
        "blk.fn.norm.gamma": 1 + 0.2 * torch.randn(1, C, 1, 1, 1),
        "blk.fn.fn.to_qkv.weight": torch.randn(384, C, 1, 1) * (2.0 / C ** 0.5),
        "blk.fn.fn.to_out.weight": torch.randn(C, 128, 1, 1) * 0.1,
        "blk.fn.fn.to_out.bias": torch.randn(C) * 0.1,
    }
    x = (torch.randn(B, C, D, H, W) * 1.5 + 0.2).half().float()
    orc = Unet3DOracle(sd)
    want = orc._spatial_linear_attn(x, "blk")
    blk = LinAttnBlock(sd["blk.fn.norm.gamma"], sd["blk.fn.fn.to_qkv.weight"], sd["blk.fn.fn.to_out.weight"],
                       sd["blk.fn.fn.to_out.bias"], device="cuda")
    xcl = x.permute(0, 2, 3, 4, 1).contiguous().half().cuda()
    got = blk(xcl).float().cpu().permute(0, 4, 1, 2, 3)
    assert rel_l2(got - x, want - x) < 6e-3, rel_l2(got - x, want - x)
    assert rel_l2(got, want) < 1e-3


@pytest.mark.parametrize("C,B,D,H,W", [(64, 1, 24, 6, 10), (128, 2, 24, 4, 5), (256, 1, 24, 3, 4), (64, 1, 7, 3, 6)])
def test_tattn_block_vs_oracle(C, B, D, H, W):
    from oracle.unet3d import Unet3DOracle, rel_pos_bias
    from wdno_b200.attn_fused import TemporalBlock
    torch.manual_seed(C + D)
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))
    sd = {
        "time_mlp.1.weight": torch.zeros(4, 1), "init_conv.weight": torch.zeros(1, 1),
        "blk.fn.norm.gamma": 1 + 0.2 * torch.randn(1, C, 1, 1, 1),
        "blk.fn.fn.fn.to_qkv.weight": torch.randn(384, C) * (2.0 / C ** 0.5),
        "blk.fn.fn.fn.to_out.weight": torch.randn(C, 128) * 0.1,
        "blk.fn.fn.fn.rotary_emb.freqs": freqs,
    }
    emb = torch.randn(32, 4) * 0.5
    pos_bias = rel_pos_bias(emb, D)
    x = (torch.randn(B, C, D, H, W) * 1.5 + 0.2).half().float()
    orc = Unet3DOracle(sd)
    want = orc._temporal_attn(x, "blk", pos_bias)
    blk = TemporalBlock(sd["blk.fn.norm.gamma"], sd["blk.fn.fn.fn.to_qkv.weight"], sd["blk.fn.fn.fn.to_out.weight"], device="cuda")
    ang = torch.arange(D, dtype=torch.float32)[:, None] * freqs[None, :]
    xcl = x.permute(0, 2, 3, 4, 1).contiguous().half().cuda()
    got = blk(xcl, bias=pos_bias.contiguous().cuda(), rot=(ang.cos().contiguous().cuda(), ang.sin().contiguous().cuda()))
    got = got.float().cpu().permute(0, 4, 1, 2, 3)
    assert rel_l2(got - x, want - x) < 6e-3, rel_l2(got - x, want - x)
    assert rel_l2(got, want) < 1e-3
