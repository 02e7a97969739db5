"""GPU probe: per-layer parity of the engine's Unet3D forward against the torch oracle + a first timing."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from oracle.unet3d import Unet3DOracle
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    B = int(os.environ.get("B", "2"))
    torch.manual_seed(0)
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if ".norm." in n or n.endswith("gamma"):
                p.add_(0.2 * torch.randn_like(p))
    m = m.cuda().eval()
    x = torch.randn(B, 24, 42, 40, 40, device="cuda")
    t = torch.randint(0, 1000, (B,), device="cuda")
    taps_e, taps_o = {}, {}
    with torch.no_grad():
        eng = m.engine()
        y = eng.forward(x, t, taps=taps_e)
        torch.cuda.synchronize()
        print("engine launches per forward:", eng.launches, flush=True)
        orc = Unet3DOracle({k: v for k, v in m.state_dict().items()})
        # oracle helpers create CPU aranges; run it on the GPU by moving its dict and patching default device
        orc.sd = {k: v.cuda() for k, v in orc.sd.items()}
        torch.set_default_device("cuda")
        yo = orc(x, t, taps=taps_o)
        torch.set_default_device("cpu")
    rows = []
    for k, vo in taps_o.items():
        if k not in taps_e:
            continue
        ve = taps_e[k].permute(0, 4, 1, 2, 3).float()
        err = float((ve - vo).abs().max() / (vo.abs().max() + 1e-9))
        rel2 = float((ve - vo).norm() / (vo.norm() + 1e-9))
        rows.append((k, err, rel2))
        print(f"{k:24s} max-rel {err:.3e}  l2-rel {rel2:.3e}", flush=True)
    err = float((y - yo).abs().max() / yo.abs().max())
    rel2 = float((y - yo).norm() / yo.norm())
    print(f"{'OUTPUT eps':24s} max-rel {err:.3e}  l2-rel {rel2:.3e}", flush=True)
    res = {"B": B, "out_maxrel": err, "out_l2rel": rel2, "layers": rows}
    # timing
    Bt = int(os.environ.get("BT", "16"))
    xt = torch.randn(Bt, 24, 42, 40, 40, device="cuda")
    tt = torch.randint(0, 1000, (Bt,), device="cuda")
    with torch.no_grad():
        for _ in range(5):
            eng.forward(xt, tt)
        torch.cuda.synchronize()
        t0 = time.time()
        while time.time() - t0 < 1.0:  # clock warm-up
            eng.forward(xt, tt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            eng.forward(xt, tt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    res["forward_ms_B%d" % Bt] = ms
    res["tflops"] = 326.35e9 * Bt / ms / 1e9
    print(f"forward B={Bt}: {ms:.2f} ms  -> {res['tflops']:.1f} TFLOP/s algorithmic", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_unet3d.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
