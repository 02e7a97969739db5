"""one training step of the smoke base model (batch N, default 6) for ncu captures: python tools/train_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wdno_b200.diffusion_smoke import GaussianDiffusion  # noqa: E402
from wdno_b200.trainer import FusedTrainer  # noqa: E402
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: E402

B = int(os.environ.get("B", "6"))
N = int(os.environ.get("N", "2"))
torch.manual_seed(0)
m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().train()
gd = GaussianDiffusion(m, torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1), True, True, True, False, "bior1.3", "zero",
                       [18, 34, 34], [32, 64, 64], image_size=40, frames=24, timesteps=1000, sampling_timesteps=250).cuda()
tr = FusedTrainer(gd, lr=1e-4)
x0 = torch.randn(B, 24, 42, 40, 40, device="cuda").clamp(-1, 1)
for i in range(N):
    loss = tr.step(x0)
torch.cuda.synchronize()
print("loss", float(loss))
