"""Which layer makes row i of a batch-16 forward differ from the batch-1 forward of the same sample?  (and: is a forward
bit-reproducible run to run?)  Prints the first layers with a non-zero difference."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wdno_b200.unet3d import Unet3D_with_Conv3D  # noqa: E402

torch.manual_seed(0)
m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().eval()
g = torch.Generator().manual_seed(3)
x = torch.randn(16, 24, 42, 40, 40, generator=g).cuda()
t = torch.randint(0, 1000, (16,), generator=g).cuda()
e = m.engine()
with torch.no_grad():
    ta, tb, t1 = {}, {}, {}
    ya = e.forward(x, t, taps=ta)
    yb = e.forward(x, t, taps=tb)
    print("run-to-run bit-equal (B=16):", bool(torch.equal(ya, yb)), [k for k in ta if not torch.equal(ta[k], tb[k])][:5])
    i = 7
    y1 = e.forward(x[i:i + 1].contiguous(), t[i:i + 1].contiguous(), taps=t1)
    for k in ta:
        a, b = ta[k][i].float(), t1[k][0].float()
        d = float((a - b).norm() / (b.norm() + 1e-30))
        nz = float((a != b).float().mean())
        print(f"{k:24s} rel {d:.3e}  differing elements {nz:.4f}")
    print("output rel", float((ya[i] - y1[0]).norm() / y1[0].norm()))
