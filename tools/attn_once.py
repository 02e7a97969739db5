"""One full-resolution fused temporal-attention block and one fused linear-attention block (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200.attn_fused import LinAttnBlock, TemporalBlock  # noqa: E402

torch.manual_seed(0)
C, B, D, H, W = 64, 16, 24, 40, 40
x = torch.randn(B, D, H, W, C, device="cuda").half()
tb = TemporalBlock(torch.ones(C), torch.randn(384, C) * 0.2, torch.randn(C, 128) * 0.1)
lb = LinAttnBlock(torch.ones(C), torch.randn(384, C, 1, 1) * 0.2, torch.randn(C, 128, 1, 1) * 0.1, torch.zeros(C))
bias = torch.zeros(4, D, D, device="cuda")
ang = torch.arange(D, dtype=torch.float32)[:, None] * (1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)))[None, :]
rot = (ang.cos().contiguous().cuda(), ang.sin().contiguous().cuda())
for _ in range(int(os.environ.get("N", "2"))):
    y = tb(x, bias=bias, rot=rot)
    z = lb(x)
torch.cuda.synchronize()
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record(); y = tb(x, bias=bias, rot=rot); e1.record(); z = lb(x); e2.record()
torch.cuda.synchronize()
print(f"tattn {e0.elapsed_time(e1)*1e3:.1f} us, linattn {e1.elapsed_time(e2)*1e3:.1f} us")
