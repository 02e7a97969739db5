"""GPU bring-up probe for the tcgen05 tap-GEMM kernel: simplest case first, each case in its own
subprocess with a timeout so a hung kernel cannot eat the whole gpurun call.

  python tools/gpu_probe_tapgemm.py            # run all cases, write gpurun_out/probe_tapgemm.log
  python tools/gpu_probe_tapgemm.py --case 3   # one case in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cl(x):
    import torch
    return x.permute(0, 2, 3, 4, 1).contiguous().to(torch.float16)


def uncl(y):
    return y.permute(0, 4, 1, 2, 3).float()


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def run_case(i):
    import torch
    import torch.nn.functional as F
    from wdno_b200.tapgemm import TapGemm
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    torch.manual_seed(100 + i)
    res = {"case": i}

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, device=dev) * scale).half().float()

    if i == 0:
        res["name"] = "1x1 16->16 single tile"
        x = rnd(1, 16, 4, 4, 32)
        w = rnd(16, 16, scale=0.2)
        plan = TapGemm(w, None, device=dev)
        out = plan(cl(x))
        ref = F.conv3d(x, w[:, :, None, None, None])
        res["err"] = relerr(uncl(out), ref)
    elif i == 1:
        res["name"] = "1x1 64->64 B=2 D=8 20x20 bias"
        x = rnd(2, 64, 8, 20, 20)
        w = rnd(64, 64, scale=0.1)
        b = rnd(64)
        plan = TapGemm(w, b, device=dev)
        out = plan(cl(x))
        ref = F.conv3d(x, w[:, :, None, None, None], b)
        res["err"] = relerr(uncl(out), ref)
    elif i == 2:
        res["name"] = "3x3x3 16->16 small"
        x = rnd(1, 16, 4, 8, 8)
        w = rnd(16, 16, 3, 3, 3, scale=0.1)
        plan = TapGemm(w, None, device=dev)
        out = plan(cl(x))
        ref = F.conv3d(x, w, padding=1)
        res["err"] = relerr(uncl(out), ref)
    elif i == 3:
        res["name"] = "3x3x3 64->64 B=2 24x40x40 bias+stats"
        x = rnd(2, 64, 24, 40, 40)
        w = rnd(64, 64, 3, 3, 3, scale=0.03)
        b = rnd(64)
        plan = TapGemm(w, b, device=dev)
        stats = torch.zeros(2, 8, 2, dtype=torch.float64, device=dev)
        out = plan(cl(x), stats=stats, groups=8)
        ref = F.conv3d(x, w, b, padding=1)
        res["err"] = relerr(uncl(out), ref)
        rs = ref.reshape(2, 8, -1).double()
        res["err_sum"] = relerr(stats[:, :, 0], rs.sum(-1))
        res["err_sumsq"] = relerr(stats[:, :, 1], (rs ** 2).sum(-1))
    elif i == 4:
        res["name"] = "3x3x3 concat(64+64)->64 affine+silu on load, B=2 24x20x20"
        x0 = rnd(2, 64, 24, 20, 20)
        x1 = rnd(2, 64, 24, 20, 20)
        a0, c0 = torch.randn(2, 64, device=dev), torch.randn(2, 64, device=dev)
        w = rnd(64, 128, 3, 3, 3, scale=0.03)
        plan = TapGemm(w, None, src_channels=(64, 64), device=dev)
        out = plan(cl(x0), cl(x1), coef0=(a0, c0))
        act = F.silu(x0 * a0[:, :, None, None, None] + c0[:, :, None, None, None]).half().float()
        ref = F.conv3d(torch.cat([act, x1], 1), w, padding=1)
        res["err"] = relerr(uncl(out), ref)
    elif i == 5:
        res["name"] = "7x7x7 init conv 42(48)->64 B=1 24x40x40"
        x = torch.zeros(1, 48, 24, 40, 40, device=dev)
        x[:, :42] = rnd(1, 42, 24, 40, 40)
        w = rnd(64, 42, 7, 7, 7, scale=0.01)
        b = rnd(64)
        plan = TapGemm(w, b, src_channels=(48,), device=dev)
        out = plan(cl(x))
        ref = F.conv3d(x[:, :42], w, b, padding=3)
        res["err"] = relerr(uncl(out), ref)
    elif i == 6:
        res["name"] = "down144 64->64 and up144 64->64, B=2 24x40x40 / 20x20"
        x = rnd(2, 64, 24, 40, 40)
        wd, bd = rnd(64, 64, 1, 4, 4, scale=0.05), rnd(64)
        plan = TapGemm(wd, bd, kind="down144", device=dev)
        out = plan(cl(x))
        ref = F.conv3d(x, wd, bd, stride=(1, 2, 2), padding=(0, 1, 1))
        res["err_down"] = relerr(uncl(out), ref)
        xs = rnd(2, 64, 24, 20, 20)
        wu, bu = rnd(64, 64, 1, 4, 4, scale=0.05), rnd(64)
        plan_u = TapGemm(wu, bu, kind="up144", device=dev)
        out_u = plan_u(cl(xs))
        ref_u = F.conv_transpose3d(xs, wu, bu, stride=(1, 2, 2), padding=(0, 1, 1))
        res["err_up"] = relerr(uncl(out_u), ref_u)
        res["err"] = max(res["err_down"], res["err_up"])
    elif i == 7:
        res["name"] = "2D mode (D=1): 3x3 128->128 at 64x64 B=2, unshuffle 128->256, up2+3x3"
        x = rnd(2, 128, 1, 64, 64)
        w, b = rnd(128, 128, 3, 3, scale=0.03), rnd(128)
        plan = TapGemm(w, b, device=dev)
        out = plan(cl(x))
        ref = F.conv2d(x[:, :, 0], w, b, padding=1)
        res["err_conv"] = relerr(uncl(out)[:, :, 0], ref)
        wq, bq = rnd(256, 512, 1, 1, scale=0.05), rnd(256)
        plan_q = TapGemm(wq, bq, kind="unshuffle", device=dev)
        out_q = plan_q(cl(x))
        xu = x[:, :, 0].reshape(2, 128, 32, 2, 32, 2).permute(0, 1, 3, 5, 2, 4).reshape(2, 512, 32, 32)
        res["err_unshuffle"] = relerr(uncl(out_q)[:, :, 0], F.conv2d(xu, wq, bq))
        xs = rnd(2, 128, 1, 32, 32)
        plan_p = TapGemm(w, b, up2=True, device=dev)
        out_p = plan_p(cl(xs))
        ref_p = F.conv2d(F.interpolate(xs[:, :, 0], scale_factor=2, mode="nearest"), w, b, padding=1)
        res["err_up2"] = relerr(uncl(out_p)[:, :, 0], ref_p)
        res["err"] = max(res["err_conv"], res["err_unshuffle"], res["err_up2"])
    elif i == 8:
        res["name"] = "N=128 tile: 3x3x3 128->256 B=2 24x10x10 + resid"
        x = rnd(2, 128, 24, 10, 10)
        w, b = rnd(256, 128, 3, 3, 3, scale=0.03), rnd(256)
        r = rnd(2, 256, 24, 10, 10)
        plan = TapGemm(w, b, device=dev, n_tile=128)
        out = plan(cl(x), resid=cl(r))
        ref = F.conv3d(x, w, b, padding=1) + r
        res["err"] = relerr(uncl(out), ref)
    elif i == 9:
        res["name"] = "final 1x1 64->42 fp32 [B,F,C,H,W] output"
        x = rnd(2, 64, 24, 40, 40)
        w, b = rnd(42, 64, scale=0.1), rnd(42)
        plan = TapGemm(w, b, device=dev)
        out = plan(cl(x), out_fp32_bfchw=True)
        ref = F.conv3d(x, w[:, :, None, None, None], b)
        res["err"] = relerr(out.permute(0, 2, 1, 3, 4), ref)
    elif i == 10:
        res["name"] = "timing: 3x3x3 64->64 B=16 24x40x40"
        x = cl(rnd(16, 64, 24, 40, 40))
        w, b = rnd(64, 64, 3, 3, 3, scale=0.03), rnd(64)
        plan = TapGemm(w, b, device=dev)
        out = plan(x)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            plan(x, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        fl = 2 * 27 * 64 * 64 * 16 * 24 * 40 * 40
        res["ms"] = ms
        res["tflops"] = fl / ms / 1e9
        res["err"] = 0.0
    elif i == 11:
        res["name"] = "timing: 3x3x3 256->256 B=16 24x10x10 (N=64 and N=128), 7^3 init B=16"
        x = cl(rnd(16, 256, 24, 10, 10))
        w, b = rnd(256, 256, 3, 3, 3, scale=0.03), rnd(256)
        for nt in (64, 128):
            plan = TapGemm(w, b, device=dev, n_tile=nt)
            out = plan(x)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                plan(x, out=out)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            res[f"ms_n{nt}"] = ms
            res[f"tflops_n{nt}"] = 2 * 27 * 256 * 256 * 16 * 2400 / ms / 1e9
        xi = torch.zeros(16, 24, 40, 40, 48, device=dev, dtype=torch.float16)
        wi, bi = rnd(64, 42, 7, 7, 7, scale=0.01), rnd(64)
        plan = TapGemm(wi, bi, src_channels=(48,), device=dev)
        out = plan(xi)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(3):
            plan(xi, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        res["ms_init7"] = ms
        res["tflops_init7"] = 2 * 343 * 42 * 64 * 16 * 38400 / ms / 1e9
        res["err"] = 0.0
    elif i == 12:
        res["name"] = "1x1 with resident slabs over N-chunks: 64->384 (24x40x40), 256->384 (24x10x10), 512->128 + resid"
        errs = []
        for cin, cout, hw in ((64, 384, 40), (256, 384, 10), (128, 256, 10)):
            x = rnd(2, cin, 24, hw, hw)
            w = rnd(cout, cin, scale=0.1)
            plan = TapGemm(w, None, device=dev)
            out = plan(cl(x))
            errs.append(relerr(uncl(out), F.conv3d(x, w[:, :, None, None, None])))
        x0, x1 = rnd(2, 256, 24, 10, 10), rnd(2, 256, 24, 10, 10)
        w, b, r = rnd(128, 512, scale=0.05), rnd(128), rnd(2, 128, 24, 10, 10)
        plan = TapGemm(w, b, src_channels=(256, 256), device=dev)
        out = plan(cl(x0), cl(x1), resid=cl(r))
        errs.append(relerr(uncl(out), F.conv3d(torch.cat([x0, x1], 1), w[:, :, None, None, None], b) + r))
        res["errs"] = errs
        res["err"] = max(errs)
    elif i == 13:
        res["name"] = "column strips: 7x7x7 82(96)->64 at 24x80x80 (planner-chosen), forced strips on 3x3x3 64->64 24x40x44 + stats, fp32 out"
        x = torch.zeros(1, 96, 24, 80, 80, device=dev)
        x[:, :82] = rnd(1, 82, 24, 80, 80)
        w, b = rnd(64, 82, 7, 7, 7, scale=0.01), rnd(64)
        plan = TapGemm(w, b, src_channels=(96,), device=dev)
        out = plan(cl(x))
        pl = plan._plan(1, 24, 80, 80)
        res["strips"], res["zstack"] = int(pl.strips), int(pl.zstack)
        ref = F.conv3d(x[:, :82], w, b, padding=3)
        errs = [relerr(uncl(out), ref)]
        if pl.strips < 2 or not pl.zstack:
            errs.append(1.0)  # the planner must choose strips + the stacked scheme here
        os.environ["WDNO_FORCE_STRIPS"] = "3"
        try:
            x = rnd(2, 64, 24, 40, 44)
            w, b = rnd(64, 64, 3, 3, 3, scale=0.03), rnd(64)
            plan = TapGemm(w, b, device=dev)
            stats = torch.zeros(2, 8, 2, dtype=torch.float64, device=dev)
            out = plan(cl(x), stats=stats, groups=8)
            ref = F.conv3d(x, w, b, padding=1)
            errs.append(relerr(uncl(out), ref))
            rs = ref.reshape(2, 8, -1).double()
            errs.append(relerr(stats[:, :, 1], (rs ** 2).sum(-1)))
            w2, b2 = rnd(42, 64, 3, 3, 3, scale=0.03), rnd(42)
            plan2 = TapGemm(w2, b2, device=dev)
            out2 = plan2(cl(x), out_fp32_bfchw=True)
            errs.append(relerr(out2.permute(0, 2, 1, 3, 4), F.conv3d(x, w2, b2, padding=1)))
            res["forced_strips"] = int(plan._plan(2, 24, 40, 44).strips)
        finally:
            del os.environ["WDNO_FORCE_STRIPS"]
        res["errs"] = errs
        res["err"] = max(errs)
    elif i == 14:
        res["name"] = "conv1x1 kernel edge cases: ragged M, concat sources, bias+resid, masked N (72, 42), planar fp32 with odd planes"
        errs = []
        x0, x1 = rnd(3, 64, 5, 7, 9), rnd(3, 96, 5, 7, 9)           # M = 945 (not a multiple of 128)
        w, b, r = rnd(72, 160, scale=0.1), rnd(72), rnd(3, 72, 5, 7, 9)
        plan = TapGemm(w, b, src_channels=(64, 96), device=dev)
        assert plan._c1 is not None
        out = plan(cl(x0), cl(x1), resid=cl(r))
        errs.append(relerr(uncl(out), F.conv3d(torch.cat([x0, x1], 1), w[:, :, None, None, None], b) + r))
        w2, b2 = rnd(42, 64, scale=0.1), rnd(42)
        plan2 = TapGemm(w2, b2, device=dev)
        out2 = plan2(cl(x0), out_fp32_bfchw=True)                    # HW = 63: every tile straddles planes
        errs.append(relerr(out2.permute(0, 2, 1, 3, 4), F.conv3d(x0, w2[:, :, None, None, None], b2)))
        x3 = rnd(2, 1536, 1, 8, 8)
        w3 = rnd(1024, 1536, scale=0.03)
        plan3 = TapGemm(w3, None, device=dev)
        errs.append(relerr(uncl(plan3(cl(x3))), F.conv3d(x3, w3[:, :, None, None, None])))
        res["errs"] = errs
        res["err"] = max(errs)
    else:
        return None
    torch.cuda.synchronize()
    return res


NCASES = 15


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=None)
    ap.add_argument("--timeout", type=int, default=120)
    ap.add_argument("--cases", type=str, default=None)
    args = ap.parse_args()
    if args.case is not None:
        r = run_case(args.case)
        print("RESULT " + json.dumps(r), flush=True)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe_tapgemm.log"), "w")
    cases = [int(c) for c in args.cases.split(",")] if args.cases else list(range(NCASES))
    for i in cases:
        t0 = time.time()
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)], capture_output=True,
                                text=True, timeout=args.timeout)
            lines = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            msg = lines[-1] if lines else f"case {i} rc={pr.returncode} NO RESULT\n{pr.stdout[-1500:]}\n{pr.stderr[-3000:]}"
        except subprocess.TimeoutExpired:
            msg = f"case {i} TIMEOUT after {args.timeout}s (hung kernel?)"
        line = f"[{time.time() - t0:6.1f}s] {msg}"
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()


if __name__ == "__main__":
    main()
