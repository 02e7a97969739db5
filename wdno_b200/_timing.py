"""Per-launch CUDA-event timing of the dominant kernels (bench.py, tools/): set `sink` to a list and every instrumented
launch appends dict(kernel, e0, e1, flops, bytes, meta) -- algorithmic FLOPs / bytes of the reference op it replaces
(SURVEY.md section 8d).  None (default) = no overhead beyond one attribute read."""
import torch

sink = None


class span:
    """with span("tattn_block", flops=..., bytes=..., meta=...): launch"""

    def __init__(self, kernel, flops=0.0, bytes=0.0, meta=None):
        self.rec = None
        if sink is not None:
            self.rec = dict(kernel=kernel, flops=float(flops), bytes=float(bytes), meta=meta,
                            e0=torch.cuda.Event(enable_timing=True), e1=torch.cuda.Event(enable_timing=True))

    def __enter__(self):
        if self.rec is not None:
            self.rec["e0"].record()
        return self

    def __exit__(self, *a):
        if self.rec is not None:
            self.rec["e1"].record()
            sink.append(self.rec)
