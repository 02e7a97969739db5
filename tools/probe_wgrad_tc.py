"""tcgen05 wgrad (csrc/wgrad_tc.cu) vs torch autograd, both descriptor-stride conventions (run under `timeout`)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wdno_b200.training import WgradTC, colsum_f16  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(1)
rnd = lambda *s: torch.randn(*s, generator=g).cuda()
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
cases = [((64, 64, 3, 3, 3), (2, 6, 12, 10)), ((128, 64, 3, 3, 3), (1, 5, 9, 11)), ((64, 48, 7, 7, 7), (1, 6, 10, 10)),
         ((256, 128, 1, 3, 3), (3, 1, 8, 8)), ((64, 64, 3, 3, 3), (1, 24, 40, 40))]
only = os.environ.get("CASE")
for swap in ([int(os.environ["SWAP"])] if "SWAP" in os.environ else [0, 1]):
    WgradTC.SWAP = swap
    for ci, (ws, (B, D, H, W)) in enumerate(cases):
        if only is not None and int(only) != ci:
            continue
        co, cin, KD, KH, KW = ws
        x = rnd(B, D, H, W, cin).half()
        dy = rnd(B, D, H, W, co).half()
        xt = x.float().permute(0, 4, 1, 2, 3)
        wt = torch.zeros(ws, device="cuda", requires_grad=True)
        yt = F.conv3d(xt, wt, None, padding=(KD // 2, KH // 2, KW // 2))
        yt.backward(dy.float().permute(0, 4, 1, 2, 3))
        dw = torch.zeros(ws, device="cuda")
        eng = WgradTC(co, KD, KH, KW, "cuda")
        print("swap", swap, "case", ci, ws, (B, D, H, W), "...", flush=True)
        eng(x, dy, dw, 1.0, cx_n=cin, n_total=cin, m_valid=co)
        torch.cuda.synchronize()
        print("   rel", rel(dw, wt.grad), "norms", float(dw.norm()), float(wt.grad.norm()), flush=True)
    db = torch.zeros(64, device="cuda")
    dyb = rnd(3, 5, 7, 9, 64).half()
    colsum_f16(dyb, db, 1.0)
    print("colsum rel", rel(db, dyb.float().sum(dim=(0, 1, 2, 3))), flush=True)
