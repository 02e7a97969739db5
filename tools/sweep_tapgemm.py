"""Tuning sweep for the tap-GEMM plan knobs on the dominant shapes (run on the GPU box)."""
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
x = torch.randn(16, 24, hw, hw, cin, device="cuda").half()
w = torch.randn(cout, cin, 3, 3, 3) * 0.03
plan = TapGemm(w, torch.randn(cout), device="cuda")
out = plan(x)
t0 = time.time()
while time.time() - t0 < 0.7:
    plan(x, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    plan(x, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
p = plan._plan(16, 24, hw, hw)
print("RES shape=%%s KC=%%d NSLOT=%%d NBST=%%d TPS=%%d N=%%d ZT=%%d : %%.3f ms  %%.0f TF/s" %% (os.environ["SHAPE"], p.KC, p.NSLOT, p.NBST, p.TPS, p.N, p.ZT, ms, 2*27*cin*cout*16*24*hw*hw/ms/1e9))
''' % ROOT
for shape in ("64,64,40", "256,256,10"):
    for kc, extra, bst in itertools.product(("64", "32", "16"), ("0", "1", "2", "4"), ("8192", "16384", "32768")):
        env = dict(os.environ, SHAPE=shape, WDNO_KC=kc, WDNO_SLOT_EXTRA=extra, WDNO_BSTAGE=bst)
        try:
            r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
            print(line[0] if line else f"shape={shape} KC={kc} extra={extra} bst={bst}: FAIL {r.stderr.strip().splitlines()[-1][:100] if r.stderr.strip() else ''}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"shape={shape} KC={kc} extra={extra} bst={bst}: TIMEOUT", flush=True)
