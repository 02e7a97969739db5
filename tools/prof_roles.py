"""Per-role stall accounting of tapgemm (debug build: WDNO_PROF=1 python -m wdno_b200.build --force).
Prints, for a few representative layers, the cycles lane 0 of each role spent waiting on each barrier."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200 import _lib  # noqa: E402
from wdno_b200.tapgemm import TapGemm  # noqa: E402

L = _lib.lib()
dev = "cuda"
torch.manual_seed(0)


def read():
    buf = (C.c_ulonglong * 32)()
    L.wdno_tapgemm_prof_read(buf)
    return list(buf)


def run(name, cin, cout, k, B, D, H, W, act=True, srcs=None, plain=False):
    w = torch.randn(cout, cin, k, k, k) * 0.05
    g = TapGemm(w, None if plain else torch.randn(cout), device=dev, src_channels=srcs)
    xs = [torch.randn(B, D, H, W, c, device=dev).half() for c in (srcs or (cin,))]
    coefs = [(torch.rand(B, c, device=dev) + 0.5, torch.randn(B, c, device=dev)) for c in (srcs or (cin,))] if act else [None, None]
    stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev)
    kw = dict(coef0=coefs[0]) if plain else dict(coef0=coefs[0], stats=stats)
    if len(xs) > 1:
        kw["coef1"] = coefs[1]
    for _ in range(3):
        g(*xs, **kw)
    read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        g(*xs, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    pr = read()
    p = g._plan(B, D, H, W)
    ctas = p.grid * n
    fl = 2.0 * B * D * H * W * cout * cin * k ** 3
    print(f"\n== {name}: {ms*1e3:.1f} us, {fl/ms/1e9:.0f} TF/s; plan KC={p.KC} ZT={p.ZT} PT={p.PT} N={p.N} NSLOT={p.NSLOT} NBST={p.NBST} TPS={p.TPS} reuse={p.reuse} grid={p.grid}")
    lab = {0: ("MMA ", ["wait acc_empty", "wait slab_full", "wait b_full", "issue", "", ""]),
           8: ("PROD", ["wait slab_empty", "issue cp.async", "cp.async wait", "transform", "", ""]),
           16: ("BW  ", ["wait b_empty", "", "", "", "", ""]),
           24: ("EPI ", ["wait acc_full", "tmem ld", "stage write", "coalesced store", "addr calc", ""])}
    for base, (role, names) in lab.items():
        tot = pr[base + 6] / ctas
        parts = ", ".join(f"{nm} {pr[base+i]/ctas/1e3:.1f}k ({100*pr[base+i]/max(1,pr[base+6]):.0f}%)" for i, nm in enumerate(names) if nm)
        print(f"  {role} total {tot/1e3:.1f}k cyc/CTA : {parts}")


run("3x3x3 64->64 @24x40x40 act", 64, 64, 3, 16, 24, 40, 40)
if len(sys.argv) > 1:
    run("1x1 64->384 @24x40x40", 64, 384, 1, 16, 24, 40, 40, act=False, plain=True)
    sys.exit(0)
run("3x3x3 64->64 @24x40x40 identity", 64, 64, 3, 16, 24, 40, 40, act=False)
run("3x3x3 256->256 @24x10x10 act", 256, 256, 3, 16, 24, 10, 10)
run("3x3x3 128->128 @24x20x20 act", 128, 128, 3, 16, 24, 20, 20)
run("3x3x3 128->64 concat @24x40x40 act", 128, 64, 3, 16, 24, 40, 40, srcs=(64, 64))
run("1x1 64->384 @24x40x40", 64, 384, 1, 16, 24, 40, 40, act=False, plain=True)
run("1x1 192->64 @24x40x40", 192, 64, 1, 16, 24, 40, 40, act=False, plain=True)
run("7x7x7 48->64 @24x40x40", 48, 64, 7, 16, 24, 40, 40, act=False)
