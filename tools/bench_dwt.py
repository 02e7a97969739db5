#!/usr/bin/env python
"""Throughput of the wavelet transforms at the BASELINE shapes (device-timed, inputs resident, L2 flushed between runs):
  3-D  wavedec3 / waverec3, bior1.3 'zero':  [16*5, 32, 64, 64] <-> 8 x [80, 18, 34, 34]      (smoke, C3/C5 batch)
  2-D  DWTForward / DWTInverse, bior2.4 'periodization': [256, 2, 81, 120] <-> [256,2,41,60] + [256,2,3,41,60]  (Burgers, C2 batch)
Reports algorithmic GB/s (5.95 MB per sample per 3-D transform, 156.5 KB per sample 2-D) against the
measured HBM peak, plus the adjoint (autograd backward) of the inverse used by guided sampling."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from wdno_b200 import wavelets as W  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6538.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6538.0
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def graph_timed(make_fn, nsets=4, reps=5):
    """GPU time per call with the Python / launch overhead removed: the call is captured once per input set into a CUDA
    graph (nsets distinct input and output buffers, together larger than the 126 MB L2) and the graph is replayed."""
    fns = [make_fn(i) for i in range(nsets)]
    keep = [f() for f in fns]  # warm-up (plans, attributes)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [f() for f in fns]
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / (reps * nsets)


def line(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"transform": name, "ms": ms, "algorithmic_MB": nbytes / 1e6, "algorithmic_GBps": gbs,
                      "frac_of_hbm_peak": gbs / peak, "peak_GBps": peak}), flush=True)


B = 16
GRAPH3D_ONLY = "--graph3d" in sys.argv  # tile sweeps: only the host-overhead-free 3-D lines
x3 = torch.randn(B * 5, 32, 64, 64, device="cuda")
wv = W.Wavelet("bior1.3")
c3 = W.wavedec3(x3, wv, mode="zero", level=1)
bytes3 = 4 * (x3.numel() + 8 * c3[0].numel())
with torch.no_grad():
  if not GRAPH3D_ONLY:
    line("wavedec3 bior1.3 zero [80,32,64,64]", timed(lambda: W.wavedec3(x3, wv, mode="zero", level=1)), bytes3)
    line("waverec3 bior1.3 zero -> [80,32,64,64]", timed(lambda: W.waverec3(c3, wv)), bytes3)
leaves = [c3[0].clone().requires_grad_()] + [v.clone().requires_grad_() for v in c3[1].values()]


def fwd_bwd():
    y = W.waverec3([leaves[0], dict(zip(c3[1].keys(), leaves[1:]))], wv)
    torch.autograd.grad(y.square().sum(), leaves)


if not GRAPH3D_ONLY:
    line("waverec3 + adjoint (guidance gradient) [80,32,64,64]", timed(fwd_bwd), 2 * bytes3)
x2 = torch.randn(256, 2, 81, 120, device="cuda")
f2, i2 = W.DWTForward(J=1, wave="bior2.4", mode="periodization"), W.DWTInverse(wave="bior2.4", mode="periodization")
yl, yh = f2(x2)
bytes2 = 4 * (x2.numel() + yl.numel() + yh[0].numel())
with torch.no_grad():
  if not GRAPH3D_ONLY:
    line("DWTForward bior2.4 per [256,2,81,120]", timed(lambda: f2(x2)), bytes2)
    line("DWTInverse bior2.4 per -> [256,2,82,120]", timed(lambda: i2((yl, yh))), bytes2)

# ---- the same transforms without host overhead (CUDA-graph replays over 4 rotating buffer sets, 380 MB > L2)
xs = [torch.randn_like(x3) for _ in range(4)]
cs = [W.wavedec3(x, wv, mode="zero", level=1) for x in xs]
with torch.no_grad():
    line("graph: wavedec3 bior1.3 zero [80,32,64,64]", graph_timed(lambda i: (lambda: W.wavedec3(xs[i], wv, mode="zero", level=1))), bytes3)
    line("graph: wavedec3_packed (builder layout) [80,32,64,64]", graph_timed(lambda i: (lambda: W.wavedec3_packed(xs[i], wv))), bytes3)
    line("graph: waverec3 bior1.3 zero -> [80,32,64,64]", graph_timed(lambda i: (lambda: W.waverec3(cs[i], wv))), bytes3)
    x5 = [x[:40] for x in xs]
    c5 = [W.wavedec3(x, wv, mode="zero", level=1) for x in x5]
    line("graph: waverec3 [40,32,64,64] (C5 per-GPU batch)", graph_timed(lambda i: (lambda: W.waverec3(c5[i], wv))), bytes3 // 2)
    line("graph: wavedec3 [40,32,64,64] (C5 adjoint shape)", graph_timed(lambda i: (lambda: W.wavedec3(x5[i], wv, mode="zero", level=1))), bytes3 // 2)
    # adjoint of waverec3 (guidance gradient): the analysis kernel with the reconstruction filters
    line("graph: waverec3 adjoint [80,32,64,64]", graph_timed(lambda i: (lambda: W._ana3d_raw(xs[i], wv.rec_lo, wv.rec_hi, 4, (18, 34, 34)))), bytes3)
    if GRAPH3D_ONLY:
        sys.exit(0)
    x2s = [torch.randn_like(x2) for _ in range(4)]
    y2s = [f2(x) for x in x2s]
    line("graph: DWTForward bior2.4 per [256,2,81,120]", graph_timed(lambda i: (lambda: f2(x2s[i]))), bytes2)
    line("graph: DWTInverse bior2.4 per -> [256,2,82,120]", graph_timed(lambda i: (lambda: i2(y2s[i]))), bytes2)
