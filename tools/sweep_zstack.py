"""Plan sweep for the kz-stacked tap-GEMM layers (64-channel 3x3x3): K-set width x weight-stage size, with / without the fused
GroupNorm prologue.  One subprocess per setting (the knobs are read when a plan is built)."""
import itertools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, time, torch
sys.path.insert(0, %r)
from wdno_b200.tapgemm import TapGemm
torch.manual_seed(0)
cin, cout, hw = [int(v) for v in os.environ["SHAPE"].split(",")]
act = os.environ["ACT"] == "1"
x = torch.randn(16, 24, hw, hw, cin, device="cuda").half()
w = torch.randn(cout, cin, 3, 3, 3) * 0.03
plan = TapGemm(w, torch.randn(cout), device="cuda")
coef = (torch.rand(16, cin, device="cuda") + 0.5, torch.randn(16, cin, device="cuda")) if act else None
stats = torch.zeros(16, 8, 2, dtype=torch.float64, device="cuda")
out = plan(x, coef0=coef, stats=stats)
t0 = time.time()
while time.time() - t0 < 0.5:
    plan(x, coef0=coef, stats=stats, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    plan(x, coef0=coef, stats=stats, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
p = plan._plan(16, 24, hw, hw)
print("RES shape=%%s act=%%d KC=%%d NSLOT=%%d NBST=%%d TPS=%%d N=%%d ZT=%%d zstack=%%d : %%.1f us  %%.0f TF/s" %% (os.environ["SHAPE"], act, p.KC, p.NSLOT, p.NBST, p.TPS, p.N, p.ZT, p.zstack, ms * 1e3, 2*27*cin*cout*16*24*hw*hw/ms/1e9))
''' % ROOT
shapes = sys.argv[1:] or ["64,64,40", "128,64,40", "64,64,20"]
for shape in shapes:
    for act in ("0", "1"):
        for kc, zst in itertools.product(("32", "16", "64"), ("40960", "24576", "12288")):
            if int(shape.split(",")[0]) % int(kc):
                continue
            env = dict(os.environ, SHAPE=shape, ACT=act, WDNO_KC=kc, WDNO_ZSTAGE=zst)
            try:
                r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=120)
                line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
                print((line[0] + f"  [ZSTAGE={zst}]") if line else f"shape={shape} act={act} KC={kc} zst={zst}: FAIL {r.stderr.strip().splitlines()[-1][:120] if r.stderr.strip() else ''}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"shape={shape} act={act} KC={kc} zst={zst}: TIMEOUT", flush=True)
