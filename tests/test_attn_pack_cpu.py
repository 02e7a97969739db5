"""Host-side operand packing of the tcgen05 attention blocks (no GPU): UMMA K-major order [K/8][rows][8], LayerNorm gain folded
into the projection weights, scale * log2(e) folded into the q rows of the temporal block; and the work order of the tap-GEMM's
CTA-pair mode (a Python restatement of csrc/tapgemm.cu::work_at)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _uncanon(t):
    """[K/8][rows][8] -> [rows][K]"""
    return t.permute(1, 0, 2).reshape(t.shape[1], -1)


def test_linattn_canon_operands_fold_the_layernorm_gain():
    from wdno_b200.attn_fused import LinAttnBlock
    C = 64
    torch.manual_seed(0)
    gamma = 1 + 0.3 * torch.randn(1, C, 1, 1, 1)
    wqkv = torch.randn(384, C, 1, 1)
    blk = LinAttnBlock.__new__(LinAttnBlock)
    blk._src = (gamma, wqkv, torch.randn(C, 128, 1, 1), torch.zeros(C))
    blk.C = C
    wq_c, wkv_c = blk._packed_canon()
    assert wq_c.shape == (C // 8, 128, 8) and wkv_c.shape == (C // 8, 256, 8) and wq_c.dtype == torch.float16
    want = (wqkv.reshape(384, C) * gamma.reshape(1, C)).half()
    assert torch.equal(_uncanon(wq_c), want[:128])
    assert torch.equal(_uncanon(wkv_c), want[128:])          # rows 0..127 = W_k (head, d), 128..255 = W_v (head, e)


def test_tattn_canon_operands_fold_gain_and_base2_scale():
    from wdno_b200.attn_fused import TemporalBlock
    C = 64
    torch.manual_seed(1)
    gamma = 1 + 0.3 * torch.randn(C)
    wqkv = torch.randn(384, C)
    wout = torch.randn(C, 128)
    blk = TemporalBlock.__new__(TemporalBlock)
    blk._src = (gamma, wqkv, wout)
    blk.C = C
    blk.scale = 32 ** -0.5
    plain_q, plain_o = blk._packed_canon()
    fold_q, fold_o = blk._packed_canon(fold_gamma=True)
    assert torch.equal(_uncanon(plain_q), wqkv.half()) and torch.equal(_uncanon(plain_o), wout.half())
    assert torch.equal(fold_o, plain_o)
    wg = wqkv * gamma.reshape(1, C)
    want = torch.cat([wg[:128] * (blk.scale * math.log2(math.e)), wg[128:]]).half()
    assert torch.equal(_uncanon(fold_q), want)


def _work_at(block, grid, i, n_work, n_chunks, cluster):
    if cluster != 2:
        w = block + i * grid
        return w if w < n_work else -1
    q = (block >> 1) + i * (grid >> 1)
    if 2 * q >= n_work:
        return -1
    return (q % n_chunks) + n_chunks * ((q // n_chunks) * 2 + (block & 1))


def test_cluster_work_order_covers_every_item_once_and_pairs_share_the_chunk():
    for units, n_chunks, grid in [(4, 1, 4), (156, 1, 148), (192, 2, 148), (10, 4, 6), (2, 2, 2)]:
        n_work = units * n_chunks
        seen = []
        for pair in range(grid // 2):
            i = 0
            while True:
                a, b = (_work_at(2 * pair + r, grid, i, n_work, n_chunks, 2) for r in (0, 1))
                assert (a < 0) == (b < 0)                   # both CTAs of a pair make the same number of trips
                if a < 0:
                    break
                assert a % n_chunks == b % n_chunks         # same N-chunk -> same weight stream, consumed in lock step
                assert b // n_chunks == a // n_chunks + 1   # neighbouring (sample, plane group, tile) units
                seen += [a, b]
                i += 1
        assert sorted(seen) == list(range(n_work))
        plain = []
        for blk in range(grid):
            i = 0
            while (w := _work_at(blk, grid, i, n_work, n_chunks, 0)) >= 0:
                plain.append(w)
                i += 1
        assert sorted(plain) == list(range(n_work))
