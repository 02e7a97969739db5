"""One tap-GEMM layer at batch 16 (for ncu captures): python tools/tapgemm_once.py cin cout hw [act]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200.tapgemm import TapGemm  # noqa: E402

cin, cout, hw = (int(v) for v in sys.argv[1:4])
act = len(sys.argv) > 4 and sys.argv[4] == "1"
torch.manual_seed(0)
x = torch.randn(16, 24, hw, hw, cin, device="cuda").half()
w = torch.randn(cout, cin, 3, 3, 3) * 0.03
plan = TapGemm(w, torch.randn(cout), device="cuda")
coef = (torch.rand(16, cin, device="cuda") + 0.5, torch.randn(16, cin, device="cuda")) if act else None
out = plan(x, coef0=coef)
for _ in range(3):
    plan(x, coef0=coef, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
