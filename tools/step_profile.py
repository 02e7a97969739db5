"""Warm, in-sequence timing of every C-ABI launch of one eager DDIM step (CUDA events around each call).
Complements the ncu launch list (cold-cache, serialised): use this for absolute per-kernel times."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench_r1_loop import B_PER_GPU, SHAPE, build_engine  # noqa: E402  (round-1 private-runner loop: same model and shapes)
from wdno_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
recs = None


class Wrap:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        if recs is None:
            return self.fn(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = self.fn(*a)
        e1.record()
        recs.append((self.name, e0, e1))
        return r


names = [n for n in dir(L) if n.startswith("wdno_")]
from wdno_b200 import _abi  # noqa: E402
for n in list(_abi.SIGNATURES) + ["wdno_tapgemm"]:
    setattr(L, n, Wrap(n, getattr(L, n)))

CONFIG = os.environ.get("CONFIG", "C3")   # C3 (default) | C4 (82-ch super model, 80x80) | C2 (Burgers Unet2D, batch 256)
if CONFIG == "C2":
    from wdno_b200.unet2d import Unet2D
    B = int(os.environ.get("B", "256"))
    torch.manual_seed(0)
    m = Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).cuda().eval()
    x = torch.randn(B, 9, 64, 64, device="cuda")
elif CONFIG == "C4":
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    B = int(os.environ.get("B", str(B_PER_GPU)))
    torch.manual_seed(0)
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=82).cuda().eval()
    x = torch.randn(B, 24, 82, 80, 80, device="cuda")
else:
    B = int(os.environ.get("B", str(B_PER_GPU)))
    m, gd = build_engine(250)
    x = torch.randn((B,) + SHAPE, device="cuda")
t = torch.full((B,), 500, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x, t)
    torch.cuda.synchronize()
    recs = []
    n = 3
    for _ in range(n):
        m(x, t)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, e0, e1 in recs:
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values()) / n
print(f"[{CONFIG}] batch {B}: sum of launches {tot:.3f} ms / forward")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} x{c // n:3d} {ms / n:8.3f} ms  {100 * ms / n / tot:5.1f}%")
