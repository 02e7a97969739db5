#!/usr/bin/env python
"""Device-timed step rates of the BASELINE.json configurations other than the headline one (bench.py = C3):

    python tools/bench_configs.py C2 C4 C5 [--steps K]

  C2  Burgers base: Unet2D(dim=128,(1,2,4,8),ch=9), DDPM-1000 ancestral steps, batch 256, one GPU
  C4  smoke super-resolution: Unet3D 82 ch on [16,24,82,80,80] per GPU (batch 128 over 8 GPUs), DDIM-250, `low` conditioning
  C5  smoke control: base model, guided DDIM-500 (reference guidance objective through the inverse DWT + adjoint),
      batch 8 per GPU (64 over 8 GPUs); guided steps run eagerly (user callback between U-Net and update)
One JSON line per configuration: steps/s, ms/step, algorithmic TFLOP/s of the whole step (SURVEY.md section 8d FLOPs per
sample x batch / step time) and its fraction of the measured sustained bf16 peak.  Not the driver's contract -- numbers
for DESIGN.md / profiles/.
"""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def peak_tf():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)).get("bf16_tflops_sustained", 1400.0) if os.path.exists(p) else 1400.0


def timed(fn, steps, warmup=5, settle=1.0):
    for _ in range(warmup):
        fn()
    t0 = time.time()
    while time.time() - t0 < settle:
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, workload, ms, flops_per_step, extra=None):
    tf = flops_per_step / (ms * 1e-3) / 1e12
    line = {"config": name, "workload": workload, "steps_per_s": 1e3 / ms, "ms_per_step": ms,
            "algorithmic_tflops_whole_step": tf, "frac_of_sustained_bf16_peak": tf / peak_tf()}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def c2(steps):
    from wdno_b200 import ops
    from wdno_b200.diffusion_burgers import GaussianDiffusion
    from wdno_b200.unet2d import Unet2D
    torch.manual_seed(0)
    B = 256
    m = Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).cuda().eval()
    gd = GaussianDiffusion(m, seq_length=(64, 64), is_wavelet=True, pad_mode="periodization", wave_type="bior2.4",
                           padded_shape=[41, 60], ori_shape=[81, 120], timesteps=1000, sampling_timesteps=1000,
                           is_condition_u0=True, is_condition_f=True).cuda()
    u0, f = torch.randn(B, 32, 64, device="cuda"), torch.randn(B, 4, 64, 64, device="cuda")
    with torch.no_grad():
        run = gd._runner("ddpm", (B, 9, 64, 64), dict(u_init=u0, f=f))
        run.x.normal_()

        def step():
            run.noise.normal_()
            run.step_graph(True)
        ms = timed(step, steps)
    report("C2", "Burgers base Unet2D(dim=128), DDPM ancestral step, batch 256, 1 GPU", ms, 56.13e9 * B,
           {"launches_per_step": m.engine().launches + 3})


def c4(steps):
    from wdno_b200 import ops
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    torch.manual_seed(0)
    B = 16
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=82).cuda().eval()
    gd = GaussianDiffusion(m, torch.ones(1), True, True, True, True, "bior1.3", "zero", [[18, 34, 34], [18, 66, 66]],
                           [[32, 64, 64], [32, 128, 128]], image_size=40, frames=24, timesteps=1000,
                           sampling_timesteps=250, ddim_sampling_eta=1.0).cuda()
    shape = (B, 24, 82, 80, 80)
    init = torch.randn(B, 24, 80, 80, device="cuda")
    control = torch.randn(B, 24, 16, 80, 80, device="cuda")
    low = torch.randn(B, 24, 40, 80, 80, device="cuda")
    with torch.no_grad():
        run = gd._runner("ddim", shape, 1, init, control, low, None)
        run.x.normal_()
        ops.apply_conditions(run.x, run.prog)

        def step():
            run.noise.normal_()
            run.step_graph(True)
        ms = timed(step, steps)
    report("C4", "smoke super-res Unet3D(82 ch) on [16,24,82,80,80] per GPU, DDIM-250 eta=1, low conditioning", ms,
           1577.39e9 * B, {"launches_per_step": m.engine().launches + 3,
                           "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30})


def c5(steps):
    from wdno_b200 import ops
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.smoke import inference_2d as inf
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    torch.manual_seed(0)
    B = 8
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().eval()
    shape_c, ori = [18, 34, 34], [32, 64, 64]
    R = torch.linspace(0.5, 3.0, 42, device="cuda").reshape(1, 1, 42, 1, 1)
    gd = GaussianDiffusion(m, R, False, True, True, False, "bior1.3", "zero", shape_c, ori, image_size=40, frames=24,
                           timesteps=1000, sampling_timesteps=500, ddim_sampling_eta=1.0, standard_fixed_ratio=100.0).cuda()
    args = types.SimpleNamespace(is_wavelet=True, wave_type="bior1.3", pad_mode="zero", is_condition_control=False,
                                 is_super_model=False, w_energy=0.0, w_init=0.1)
    design_fn = inf.make_design_fn(args, shape_c, ori, R)
    init = torch.randn(B, 24, 40, 40, device="cuda")
    init_u = torch.randn(B, 64, 64, device="cuda")
    design, gs = gd._guidance(design_fn, "standard", None, init, init_u)
    shape = (B, 24, 42, 40, 40)
    with torch.no_grad():
        run = gd._runner("ddim", shape, 0, init, None, None, gs)
        run.x.normal_()
        ops.apply_conditions(run.x, run.prog)

        def step():
            run.noise.normal_()
            run.step_guided(True, design)
        ms = timed(step, steps)

        def step_plain():
            run.noise.normal_()
            run.step_eager(True)
        ms_plain = timed(step_plain, steps)

        def step_gg():
            run.noise.normal_()
            run.step_guided(True, design, graph=True)
        ms_gg = timed(step_gg, steps)

        def step_graph():
            run.noise.normal_()
            run.step_graph(True)
        ms_graph = timed(step_graph, steps)
        extra_cf = {}
        try:  # closed-form guidance gradient (no autograd; opt-in, see smoke/inference_2d.py::guidance_fn_closed_form)
            design_cf, _ = gd._guidance(inf.make_design_fn(args, shape_c, ori, R, closed_form=True), "standard", None, init, init_u)

            def step_cf():
                run.noise.normal_()
                run.step_guided(True, design_cf)

            def step_cf_gg():
                run.noise.normal_()
                run.step_guided(True, design_cf, graph=True)
            extra_cf = {"ms_per_step_closed_form_guidance": timed(step_cf, steps),
                        "ms_per_step_closed_form_guidance_whole_step_graph": timed(step_cf_gg, steps)}
        except Exception as e:  # noqa: BLE001 - report, do not lose the other numbers
            extra_cf = {"closed_form_guidance_error": f"{type(e).__name__}: {e}"}
    report("C5", "smoke control: base Unet3D, guided DDIM-500 (inverse DWT + adjoint per step), batch 8 per GPU", ms,
           326.35e9 * B, {"ms_per_step_unguided_eager": ms_plain, "guidance_overhead_ms": ms - ms_plain,
            "ms_per_step_whole_step_graph(graph_design_fn=True)": ms_gg, "steps_per_s_whole_step_graph": 1e3 / ms_gg,
            "ms_per_step_unguided_graph": ms_graph, "guidance_overhead_ms_whole_step_graph": ms_gg - ms_graph, **extra_cf})


def t3(steps):
    """training step of the smoke base model at the reference's batch size (train_2d.py:37: batch 6): forward + backward on the
    engine + fused clip/Adam/EMA.  FLOPs: 3 x the forward contractions (dgrad + wgrad), SURVEY.md section 8 row a8."""
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.trainer import FusedTrainer
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    torch.manual_seed(0)
    B = 6
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).cuda().train()
    gd = GaussianDiffusion(m, torch.linspace(0.5, 3.0, 42).reshape(1, 1, 42, 1, 1), True, True, True, False, "bior1.3", "zero",
                           [18, 34, 34], [32, 64, 64], image_size=40, frames=24, timesteps=1000, sampling_timesteps=250).cuda()
    tr = FusedTrainer(gd, lr=1e-4)
    x0 = torch.randn(B, 24, 42, 40, 40, device="cuda").clamp(-1, 1)
    ms = timed(lambda: tr.step(x0), steps, warmup=3, settle=1.0)
    # split: forward only (no grad), forward + backward, optimiser
    with torch.no_grad():
        t = torch.randint(0, 1000, (B,), device="cuda")
        ms_fwd = timed(lambda: m(x0, t), steps, warmup=2, settle=0.3)

    def fb():
        tr.g.zero_()
        tr.backward(gd(x0))
    ms_fb = timed(fb, steps, warmup=2, settle=0.3)
    ms_opt = timed(tr.optimizer_step, steps, warmup=2, settle=0.2)
    # where the backward goes: CUDA events around every tape record of one step
    te = m._train_engine
    te.profile = []
    fb()
    torch.cuda.synchronize()
    agg = {}
    for kind, ch, grid, a, b in te.profile:
        k = f"{kind}:{ch}:{'x'.join(map(str, grid))}"
        agg[k] = agg.get(k, 0.0) + a.elapsed_time(b)
    te.profile = None
    by_kind = {}
    for k, v in agg.items():
        by_kind[k.split(":")[0]] = by_kind.get(k.split(":")[0], 0.0) + v
    print(json.dumps({"T3_backward_ms_by_kind": by_kind, "T3_backward_ms_by_record": dict(sorted(agg.items(), key=lambda kv: -kv[1])[:12])}), flush=True)
    report("T3", "smoke base training step (p_losses fwd + engine backward + fused clip/Adam/EMA), batch 6, 1 GPU", ms,
           3 * 326.35e9 * B, {"samples_per_s": B * 1e3 / ms, "ms_forward_only": ms_fwd, "ms_forward_backward": ms_fb,
                              "ms_optimizer": ms_opt, "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C2", "C4", "C5"])
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    for c in a.configs:
        {"C2": c2, "C4": c4, "C5": c5, "T3": t3}[c](a.steps)
