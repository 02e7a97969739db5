"""Plan sweep for the generic (N = 128) tap-GEMM layers: fewer slab slots / narrower K-sets in exchange for more taps per weight
stage (fewer tcgen05.commit per MMA: tools/micro/mma_commit_cost.cu measures ~67 issue cycles per commit)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "tools", "sweep_zstack.py")).read()
CODE = src[src.index("CODE = r'''") + len("CODE = r'''"):src.index("''' % ROOT")] % ROOT
SETTINGS = [{}, {"WDNO_SLOT_EXTRA": "1"}, {"WDNO_SLOT_EXTRA": "0"}, {"WDNO_KC": "32"}, {"WDNO_KC": "32", "WDNO_SLOT_EXTRA": "1"},
            {"WDNO_KC": "32", "WDNO_SLOT_EXTRA": "0"}, {"WDNO_KC": "32", "WDNO_BSTAGE": "73728"}, {"WDNO_BSTAGE": "32768"},
            {"WDNO_ZT": "4"}, {"WDNO_ZT": "4", "WDNO_KC": "32"}, {"WDNO_ZT": "2", "WDNO_KC": "32", "WDNO_SLOT_EXTRA": "0"}]
for shape in (sys.argv[1:] or ["256,256,10", "128,128,20", "128,256,10"]):
    for st in SETTINGS:
        r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, SHAPE=shape, ACT="1", **st), capture_output=True, text=True, timeout=120)
        line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
        print((line[0] if line else "FAIL " + (r.stderr.strip().splitlines() or ["?"])[-1][:100]), st, flush=True)
