"""Per-shape timing of every tap-GEMM launch in one U-Net forward (CUDA events, eager launches, warm clocks).
CONFIG=C3 (default) | C4 (82-ch super model on 80x80) | C2 (Burgers Unet2D, batch 256)"""
import collections
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200.tapgemm import TapGemm  # noqa: E402
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: E402

CONFIG = os.environ.get("CONFIG", "C3")
torch.manual_seed(0)
if CONFIG == "C2":
    from wdno_b200.unet2d import Unet2D
    B = int(os.environ.get("B", "256"))
    m = Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).cuda().eval()
    x = torch.randn(B, 9, 64, 64, device="cuda")
elif CONFIG == "C4":
    B = int(os.environ.get("B", "16"))
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=82).cuda().eval()
    x = torch.randn(B, 24, 82, 80, 80, device="cuda")
else:
    B = int(os.environ.get("B", "16"))
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().eval()
    x = torch.randn(B, 24, 42, 40, 40, device="cuda")
t = torch.randint(0, 1000, (B,), device="cuda")
with torch.no_grad():
    t0 = time.time()
    while time.time() - t0 < 1.5:
        m(x, t)
    torch.cuda.synchronize()
    TapGemm.timing = []
    n = 3
    for _ in range(n):
        m(x, t)
    torch.cuda.synchronize()
recs = TapGemm.timing
TapGemm.timing = None
agg = collections.OrderedDict()
for a, b, f, shp in recs:
    d = agg.setdefault(shp, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
    d[2] += f
tot = sum(v[1] for v in agg.values()) / n
print(f"[{CONFIG}] batch {B}: total tapgemm ms/forward {tot:.2f}")
rows = []
for shp, (c, ms, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rows.append(dict(shape=shp, count=c // n, ms_total=ms / n, ms_each=ms / c, tflops=f / ms / 1e9))
    print(f"{str(shp):64s} x{c // n:2d}  {ms / n:7.3f} ms  each {ms / c:6.3f}  {f / ms / 1e9:7.1f} TF/s")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"tapgemm_breakdown_{CONFIG}.json"), "w"), indent=1)
