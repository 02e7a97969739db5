"""Run N Unet3D forwards at batch B (for ncu launch lists / captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: E402

B = int(os.environ.get("B", "16"))
N = int(os.environ.get("N", "2"))
torch.manual_seed(0)
if os.environ.get("CONFIG") == "C4":   # super-resolution model: 82 channels on 80x80
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=82).cuda().eval()
    x = torch.randn(B, 24, 82, 80, 80, device="cuda")
else:
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().eval()
    x = torch.randn(B, 24, 42, 40, 40, device="cuda")
t = torch.randint(0, 1000, (B,), device="cuda")
with torch.no_grad():
    for _ in range(N):
        y = m(x, t)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()), "launches", m.engine().launches)
