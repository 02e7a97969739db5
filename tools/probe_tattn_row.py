"""all-tcgen05 temporal attention block (csrc/tattn_row.cu) against the mma.sync kernel and the first tcgen05 form; timings.
    python tools/probe_tattn_row.py [case ...]      (run under `timeout`: a wrong barrier protocol hangs)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200.attn_fused import TemporalBlock  # noqa: E402

CASES = {"tiny": (1, 7, 3, 6), "small": (1, 24, 6, 10), "odd": (2, 24, 7, 9), "mid": (2, 24, 40, 40), "full": (16, 24, 40, 40), "f32": (1, 32, 5, 8)}


def run(name):
    B, D, H, W = CASES[name]
    C = 64
    torch.manual_seed(len(name) + D)
    gamma = 1 + 0.2 * torch.randn(C)
    wqkv = torch.randn(384, C) * (2.0 / C ** 0.5)
    wout = torch.randn(C, 128) * 0.1
    x = (torch.randn(B, D, H, W, C) * 1.5 + 0.2).half().cuda()
    bias = (torch.randn(4, D, D) * 0.5).cuda()
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))
    ang = torch.arange(D, dtype=torch.float32)[:, None] * freqs[None, :]
    rot = (ang.cos().contiguous().cuda(), ang.sin().contiguous().cuda())
    blocks = {}
    for mode in ("2", "1", "0"):
        os.environ["WDNO_TATTN_TC"] = mode
        blocks[mode] = TemporalBlock(gamma, wqkv, wout, device="cuda")
    assert blocks["2"].row and blocks["1"].tc and not blocks["0"].tc and not blocks["0"].row
    ys = {}
    for mode in ("0", "1", "2"):
        ys[mode] = blocks[mode](x, bias=bias, rot=rot).float()
        torch.cuda.synchronize()
    xf = x.float()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    line = {"case": name, "shape": (B, D, H, W), "row_vs_mma_branch": rel(ys["2"] - xf, ys["0"] - xf), "tc_vs_mma_branch": rel(ys["1"] - xf, ys["0"] - xf),
            "nan": bool(torch.isnan(ys["2"]).any())}

    def timed(blk):
        for _ in range(3):
            blk(x, bias=bias, rot=rot)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            blk(x, bias=bias, rot=rot)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 100.0
    for mode, key in (("2", "us_row"), ("1", "us_tc"), ("0", "us_mma")):
        line[key] = timed(blocks[mode])
    print(line, flush=True)


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run(c)
