"""tcgen05 linear-attention block (csrc/linattn_tc.cu) against the mma.sync form and an fp32 torch restatement; timings of both.
    python tools/probe_linattn_tc.py [case ...]      (run under `timeout`: a wrong barrier protocol hangs)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200.attn_fused import LinAttnBlock  # noqa: E402

CASES = {
    "small64": (64, 6, 1600), "small128": (128, 5, 400), "tiny": (64, 2, 63), "split64": (64, 200, 300),
    "split128": (128, 333, 400), "one": (64, 1, 128), "ragged": (64, 151, 1000),
    "full64": (64, 384, 1600), "full128": (128, 384, 400),
}


def torch_ref(x, gamma, wqkv, wout, bout, eps=1e-5):
    """x [n_img, n, C] fp32"""
    mean = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    xn = (x - mean) / (var + eps).sqrt() * gamma
    qkv = xn @ wqkv.t()                                   # [I, n, 384]
    q, k, v = [t.reshape(t.shape[0], t.shape[1], 4, 32) for t in qkv.chunk(3, dim=-1)]
    q = q.softmax(dim=-1) * 32 ** -0.5
    k = k.softmax(dim=1)
    ctx = torch.einsum("inhd,inhe->ihde", k, v)
    out = torch.einsum("ihde,inhd->inhe", ctx, q).reshape(x.shape[0], x.shape[1], 128)
    return x + out @ wout.t() + bout


def run(name):
    C, n_img, n = CASES[name]
    torch.manual_seed(hash(name) % 1000)
    gamma = 1 + 0.2 * torch.randn(C)
    wqkv = torch.randn(384, C) * (2.0 / C ** 0.5)
    wout = torch.randn(C, 128) * 0.1
    bout = torch.randn(C) * 0.1
    x = (torch.randn(n_img, 1, n, 1, C) * 1.5 + 0.2).half().cuda()
    os.environ["WDNO_LINATTN_TC"] = "force"
    tc = LinAttnBlock(gamma, wqkv.reshape(384, C, 1, 1), wout.reshape(C, 128, 1, 1), bout, device="cuda")
    os.environ["WDNO_LINATTN_TC"] = "0"
    old = LinAttnBlock(gamma, wqkv.reshape(384, C, 1, 1), wout.reshape(C, 128, 1, 1), bout, device="cuda")
    assert tc.tc and not old.tc
    y_tc = tc(x)
    torch.cuda.synchronize()
    y_old = old(x)
    torch.cuda.synchronize()
    xf = x.float().reshape(n_img, n, C)
    ref = torch_ref(xf.double(), gamma.cuda().double(), wqkv.cuda().double(), wout.cuda().double(), bout.cuda().double()).float()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    yt, yo = y_tc.float().reshape(n_img, n, C), y_old.float().reshape(n_img, n, C)
    line = {"case": name, "C": C, "n_img": n_img, "n": n, "tc_branch_vs_ref": rel(yt - xf, ref - xf), "old_branch_vs_ref": rel(yo - xf, ref - xf),
            "tc_vs_ref": rel(yt, ref), "nan": bool(torch.isnan(yt).any())}
    def timed(blk):
        for _ in range(3):
            blk(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            blk(x)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 100.0
    line["us_tc"] = timed(tc)
    line["us_old"] = timed(old)
    print(line, flush=True)


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run(c)
