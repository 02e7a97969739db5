"""One small parity case per process for tests/test_gpu_knobs.py: the WDNO_* tuning knobs are read when a plan is built or a
kernel first launches, so every setting needs a fresh interpreter.  Unet3D forward vs the reference golden, a batch-8 Unet2D
forward (batch folding) and a 3-D wavelet round trip vs the fp32 / float64 oracles."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


from wdno_b200.unet2d import Unet2D  # noqa: E402
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: E402
from wdno_b200 import wavelets as W  # noqa: E402
from oracle.unet2d import Unet2DOracle  # noqa: E402
from oracle import wavelets as OW  # noqa: E402

out = {}
gold = torch.load(os.path.join(ROOT, "tests", "golden", "smoke_unet3d_fwd.pt"))
torch.manual_seed(0)
m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().eval()
x = torch.randn(1, 24, 42, 40, 40, generator=torch.Generator().manual_seed(1))
with torch.no_grad():
    y = m(x.cuda(), gold["t"].cuda())
out["unet3d"] = rel(y.reshape(-1)[::gold["stride"]].cpu(), gold["y_sub"])
torch.manual_seed(0)
m2 = Unet2D(dim=32, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).cuda().eval()
x2 = torch.randn(8, 9, 64, 64, generator=torch.Generator().manual_seed(2)).cuda()
t2 = torch.randint(0, 1000, (8,), generator=torch.Generator().manual_seed(3)).cuda()
with torch.no_grad():
    y2 = m2(x2, t2)
    yo = Unet2DOracle({k: v.detach() for k, v in m2.state_dict().items()})(x2, t2)
out["unet2d"] = rel(y2, yo)
xw = torch.randn(3, 32, 64, 64, generator=torch.Generator().manual_seed(4))
aaa, d = W.wavedec3(xw.cuda(), "bior1.3")
aaa_o, d_o = OW.wavedec3(xw.double().numpy(), "bior1.3")
out["wavedec3"] = max(float((d[k].cpu().double() - torch.from_numpy(d_o[k])).abs().max()) for k in d)
out["waverec3"] = float((W.waverec3([aaa, d], "bior1.3").cpu() - xw).abs().max())
print("KNOB_JSON " + json.dumps(out), flush=True)
