"""tap-GEMM CTA-pair multicast (p.cluster = 2) against independent CTAs on the same layer: where do the outputs differ?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from wdno_b200.tapgemm import TapGemm  # noqa: E402

CASES = {"c0": (1, 16, 16, 4, 4, 32, 1), "c2": (1, 16, 16, 4, 8, 8, 3), "c3": (2, 64, 64, 24, 40, 40, 3), "c256": (16, 256, 256, 24, 10, 10, 3)}
name = sys.argv[1]
B, Ci, Co, D, H, W, k = CASES[name]
torch.manual_seed(0)
x = (torch.randn(B, Ci, D, H, W, device="cuda")).half().float()
w = (torch.randn(Co, Ci, *([k, k, k] if k > 1 else [])) * 0.1).half().float().cuda()
xcl = x.permute(0, 2, 3, 4, 1).contiguous().half()
ref = F.conv3d(x, w if k > 1 else w[:, :, None, None, None], padding=k // 2)
outs = {}
for mode in ("0", "1"):
    os.environ["WDNO_CLUSTER"] = mode
    plan = TapGemm(w, None, device="cuda")
    out = plan(xcl)
    torch.cuda.synchronize()
    pp = plan._plan(B, D, H, W)
    print("mode", mode, "c1", plan._c1 is not None, "cluster", pp.cluster, "grid", pp.grid, "ZT", pp.ZT, "PT", pp.PT, "n_chunks", pp.n_chunks, "NBST", pp.NBST,
          "TPS", pp.TPS, "zstack", pp.zstack, "KC", pp.KC, flush=True)
    o = out.float().permute(0, 4, 1, 2, 3)
    outs[mode] = o
    err = (o - ref).abs()
    print("mode", mode, "rel err", float(err.norm() / ref.norm()))
    if mode != "0":
        per = err.amax(dim=(1, 3, 4))       # [B, D]
        print("max abs err per (b, z):", per.cpu().numpy().round(3).tolist()[:4])
        perc = err.amax(dim=(0, 2, 3, 4))
        print("max abs err per channel block of 16:", perc.reshape(-1, 16).amax(1).cpu().numpy().round(3).tolist()[:16])
    # timing
    for _ in range(3):
        plan(xcl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan(xcl)
    e1.record()
    torch.cuda.synchronize()
    print("mode", mode, "us", e0.elapsed_time(e1) * 100)
