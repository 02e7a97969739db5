"""Run a few CUDA-graph-replayed DDIM steps of the C3 workload (for ncu launch lists / full captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench_r1_loop import B_PER_GPU, SHAPE, build_engine  # noqa: E402  (round-1 private-runner loop: same model and shapes)
from wdno_b200 import ops  # noqa: E402

N = int(os.environ.get("N", "2"))
B = int(os.environ.get("B", str(B_PER_GPU)))
m, gd = build_engine(250)
shape = (B,) + SHAPE
g = torch.Generator().manual_seed(1234)
init = torch.randn(B, 24, 40, 40, generator=g).cuda()
control = torch.randn(B, 24, 16, 40, 40, generator=g).cuda()
with torch.no_grad():
    run = gd._runner("ddim", shape, 0, init, control, None, None)
    run.x.normal_()
    ops.apply_conditions(run.x, run.prog)
    for _ in range(N):
        run.noise.normal_()
        run.step_graph(True)
torch.cuda.synchronize()
print("ok", float(run.x.abs().mean()), "launches/forward", m.engine().launches)
