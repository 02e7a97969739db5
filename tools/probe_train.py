"""stage-by-stage probe of the training kernels (each stage prints before / after; run under `timeout`)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
stage = sys.argv[1]


def say(*a):
    print(*a, flush=True)


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
if stage == "layers":
    import torch.nn.functional as F
    from wdno_b200.tapgemm import TapGemm
    from wdno_b200.training import ConvLayer
    g = torch.Generator().manual_seed(1)
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()
    rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    cases = [("conv", (64, 64, 3, 3, 3), (64,), (2, 6, 12, 10)), ("conv", (64, 128, 3, 3, 3), (64, 64), (1, 5, 9, 11)),
             ("conv", (128, 64, 1, 1, 1), (64,), (2, 4, 8, 8)), ("conv", (64, 48, 7, 7, 7), (48,), (1, 6, 10, 10)),
             ("down144", (64, 64, 1, 4, 4), (64,), (2, 3, 6, 8)), ("up144", (64, 64, 1, 4, 4), (64,), (2, 3, 6, 8))]
    only = int(sys.argv[2]) if len(sys.argv) > 2 else None
    for ci, (kind, ws, srcs, (B, D, H, W)) in enumerate(cases):
        if only is not None and ci != only:
            continue
        say("case", ci, kind, ws)
        weight = torch.nn.Parameter(0.05 * rnd(*ws))
        bias = torch.nn.Parameter(0.1 * rnd(ws[1] if kind == "up144" else ws[0]))
        weight.grad, bias.grad = torch.zeros_like(weight), torch.zeros_like(bias)
        fwd = TapGemm(weight, bias, kind=kind, src_channels=srcs if kind == "conv" else None, device="cuda")
        layer = ConvLayer(fwd, weight, bias, kind, srcs)
        Hs, Ws = (2 * H, 2 * W) if kind == "down144" else (H, W)
        x = rnd(B, D, Hs, Ws, sum(srcs)).half()
        xs = [t.contiguous() for t in x.split(list(srcs), dim=-1)]
        xt = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
        wt, bt = weight.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
        if kind == "conv":
            yt = F.conv3d(xt, wt, bt, padding=(ws[2] // 2, ws[3] // 2, ws[4] // 2))
        elif kind == "down144":
            yt = F.conv3d(xt, wt, bt, stride=(1, 2, 2), padding=(0, 1, 1))
        else:
            yt = F.conv_transpose3d(xt, wt, bt, stride=(1, 2, 2), padding=(0, 1, 1))
        dy = rnd(*yt.shape).half()
        yt.backward(dy.float())
        dy_cl = dy.permute(0, 2, 3, 4, 1).contiguous()
        say("  wgrad ...")
        if kind == "up144":
            layer.backward_weight((dy_cl,), xs[0], 1.0)
        else:
            layer.backward_weight(tuple(xs), dy_cl, 1.0)
        torch.cuda.synchronize()
        say("  wgrad rel", rel(weight.grad, wt.grad), "bias", rel(bias.grad, bt.grad))
        off = 0
        for i, cs in enumerate(srcs):
            say("  dgrad", i, "...")
            dx = layer.backward_input(dy_cl, i)
            torch.cuda.synchronize()
            say("  dgrad rel", rel(dx.float(), xt.grad[:, off:off + cs].permute(0, 2, 3, 4, 1)))
            off += cs
elif stage == "norms":
    import test_gpu_training as T
    T.test_groupnorm_silu_and_layernorm_backward_vs_torch_autograd()
    say("norms ok")
elif stage == "golden":
    import test_gpu_training as T
    from wdno_b200 import train3d
    orig = train3d.Unet3DTrainEngine.backward

    # trace every tape record
    def traced(self, d_eps):
        tape = self.tape
        say("tape records", len(tape), [r[0] for r in tape][:8], "...")
        return orig(self, d_eps)
    train3d.Unet3DTrainEngine.backward = traced
    _acc = train3d.add_f16

    t0 = time.time()
    try:
        T.test_p_losses_backward_reproduces_reference_golden()
        say("golden ok", time.time() - t0)
    except AssertionError as e:
        say("golden assertion", str(e)[:2000])
elif stage == "twostep":
    import test_gpu_training as T
    T.test_two_optimizer_steps_track_the_fp32_oracle()
    say("two-step ok")
