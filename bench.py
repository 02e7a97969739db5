#!/usr/bin/env python
"""Headline benchmark: DDIM steps/s of the 2D-smoke base model (Unet3D_with_Conv3D, 24x42x40x40 wavelet coefficients,
DDIM-250, eta=1, batch 16 per GPU) -- BASELINE.json configs[2] ("C3" in SURVEY.md section 8d).

    python bench.py --gpus N --steps K --warmup W [--config C3|C4|C5|C2]   # engine arm (torchrun: one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W ...     # CPU arm: the reference's PyTorch path on the host cores

One "step" = one full-batch U-Net forward + fused DDIM update + condition re-imposition + that step's noise draw.
Every leg of the engine arm goes through the PRODUCT path `wdno_b200.parallel.sample_sharded(diffusion, batch, post=...)`
(= `GaussianDiffusion.sample()` on this rank's batch shard -> inverse transforms to fields -> one all-gather):
  value : K-step chain, conditions resident in HBM, timed on the device (CUDA events, max over ranks);
  e2e   : the configuration's full chain (DDIM-250 for C3) with pinned HOST conditions in and the gathered FIELDS copied
          back to pinned host memory inside the timed region;
  roofline / kernels : CUDA events around every launch of the dominant kernels over eager steps.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DDIM steps/sec, 2D smoke Unet3D wavelet (batch 16 per GPU per step)"   # same string on both arms (config C3)

CONFIGS = {
    # name: per-GPU batch, state shape (F, C, H, W), contraction FLOPs per sample per forward, chain length
    "C3": dict(b=16, shape=(24, 42, 40, 40), flops=326.35e9, chain=250, kind="ddim", metric=METRIC,
               workload="smoke base-res sim: Unet3D_with_Conv3D(dim=64,(1,2,4),ch=42), state [16,24,42,40,40]/GPU, DDIM-250 eta=1"),
    "C4": dict(b=16, shape=(24, 82, 80, 80), flops=1577.39e9, chain=250, kind="ddim",
               metric="DDIM steps/sec, 2D smoke super-resolution Unet3D wavelet (batch 16 per GPU per step)",
               workload="smoke super-res sim: Unet3D_with_Conv3D(ch=82), state [16,24,82,80,80]/GPU (batch 128 over 8), DDIM-250 eta=1, low conditioning"),
    "C5": dict(b=8, shape=(24, 42, 40, 40), flops=326.35e9, chain=500, kind="ddim",
               metric="guided DDIM steps/sec, 2D smoke control Unet3D wavelet (batch 8 per GPU per step)",
               workload="smoke control: base Unet3D, guided DDIM-500 (stock objective through the inverse DWT + adjoint every step), "
                        "state [8,24,42,40,40]/GPU (batch 64 over 8)"),
    "C2": dict(b=256, shape=(9, 64, 64), flops=56.13e9, chain=1000, kind="ddpm",
               metric="DDPM steps/sec, 1D Burgers Unet2D wavelet (batch 256 per GPU per step)",
               workload="Burgers base-res sim: Unet2D(dim=128,(1,2,4,8),ch=9), state [256,9,64,64]/GPU, DDPM-1000 ancestral"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(sustained=d.get("bf16_tflops_sustained"), burst=d.get("bf16_tflops"), hbm=d.get("hbm_gbs", 6650.0),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(sustained=1400.0, burst=1670.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    def __init__(self, index, period=0.02):
        self.rows, self.stop = [], threading.Event()
        self.index, self.period = index, period
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            self.stop.wait(self.period)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.strip().lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# =============================================================================== workloads (engine arm)
def host_conditions(cfg_name, batch, seed=1234):
    """synthetic conditions of the whole job (`batch` trajectories) on the host; same generator on every rank"""
    import torch
    g = torch.Generator().manual_seed(seed)
    if cfg_name == "C3":
        return dict(init=torch.randn(batch, 24, 40, 40, generator=g), control=torch.randn(batch, 24, 16, 40, 40, generator=g))
    if cfg_name == "C4":
        return dict(init=torch.randn(batch, 24, 80, 80, generator=g), control=torch.randn(batch, 24, 16, 80, 80, generator=g),
                    low=torch.randn(batch, 24, 40, 80, 80, generator=g))
    if cfg_name == "C5":
        return dict(init=torch.randn(batch, 24, 40, 40, generator=g), init_u=torch.randn(batch, 64, 64, generator=g))
    if cfg_name == "C2":
        return dict(u_init=torch.randn(batch, 32, 64, generator=g), f=torch.randn(batch, 4, 64, 64, generator=g))
    raise ValueError(cfg_name)


def build_engine(cfg_name, dev):
    """-> (model, diffusion, post(state)->fields, extra sample kwargs).  Modules and arguments are the reference's own
    (smoke/ddpm/utils.py:92-134, smoke/inference_2d.py:69-96, burgers/train_ddpm_burgers.py:149-181); random init, seed 0."""
    import types

    import torch
    torch.manual_seed(0)
    if cfg_name == "C2":
        from wdno_b200.diffusion_burgers import GaussianDiffusion
        from wdno_b200.unet2d import Unet2D
        from wdno_b200.burgers.eval_glue import coef_state_to_trajectory
        m = Unet2D(dim=128, dim_mults=[1, 2, 4, 8], channels=9, out_dim=9, resnet_block_groups=1).to(dev).eval()
        gd = GaussianDiffusion(m, seq_length=(64, 64), is_wavelet=True, pad_mode="periodization", wave_type="bior2.4",
                               padded_shape=[41, 60], ori_shape=[81, 120], timesteps=1000, sampling_timesteps=1000,
                               is_condition_u0=True, is_condition_f=True).to(dev)
        post = lambda x: coef_state_to_trajectory(x, [41, 60], [81, 120], "bior2.4", "periodization")
        return m, gd, post, {}
    from wdno_b200.diffusion_smoke import GaussianDiffusion
    from wdno_b200.smoke import inference_2d as inf
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    ch = 82 if cfg_name == "C4" else 42
    m = Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=ch).to(dev).eval()
    R = torch.linspace(0.5, 3.0, ch, device=dev).reshape(1, 1, ch, 1, 1)     # per-channel RESCALER (data_2d.py)
    if cfg_name == "C4":
        shape, ori = [[18, 34, 34], [18, 66, 66]], [[32, 64, 64], [32, 128, 128]]
        gd = GaussianDiffusion(m, R, True, True, True, True, "bior1.3", "zero", shape, ori, image_size=40, frames=24,
                               timesteps=1000, sampling_timesteps=250, ddim_sampling_eta=1.0).to(dev)
        post = lambda x: inf.state_to_fields(x, R, shape[1], ori[1], "bior1.3", "zero", "space")
        return m, gd, post, dict(N_upsample=1)
    shape, ori = [18, 34, 34], [32, 64, 64]
    control = cfg_name == "C3"
    gd = GaussianDiffusion(m, R, control, True, True, False, "bior1.3", "zero", shape, ori, image_size=40, frames=24,
                           timesteps=1000, sampling_timesteps=250 if control else 500, ddim_sampling_eta=1.0,
                           standard_fixed_ratio=100.0).to(dev)
    post = lambda x: inf.state_to_fields(x, R, shape, ori, "bior1.3", "zero")
    extra = {}
    if cfg_name == "C5":
        args = types.SimpleNamespace(is_wavelet=True, wave_type="bior1.3", pad_mode="zero", is_condition_control=False,
                                     is_super_model=False, w_energy=0.0, w_init=0.1)
        extra = dict(design_fn=inf.make_design_fn(args, shape, ori, R), design_guidance="standard")
        gd.graph_design_fn = True   # the stock objective is pure device work: whole guided step from one CUDA graph
    return m, gd, post, extra


# =============================================================================== reference arm (CPU)
def _cpu_models(cfg_name):
    """-> (kind, sample_fn(steps) -> seconds per step list).  The real reference modules when /root/reference is mounted
    (build container), else the oracle port (GPU box)."""
    import torch
    from oracle import ref_loader
    assert cfg_name == "C3", "the CPU arm times the headline configuration"
    B = CONFIGS["C3"]["b"]
    g = torch.Generator().manual_seed(1234)
    init = torch.randn(B, 24, 40, 40, generator=g)
    control = torch.randn(B, 24, 16, 40, 40, generator=g)
    if ref_loader.available():
        s = ref_loader.smoke()
        torch.manual_seed(0)
        m = s.Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).eval()
        gd = s.GaussianDiffusion(m, torch.ones(1), True, True, True, False, "bior1.3", "zero", [18, 34, 34], [32, 64, 64],
                                 image_size=40, frames=24, timesteps=1000, sampling_timesteps=250, ddim_sampling_eta=1.0)

        def run(b, steps):
            """the reference's own public API: GaussianDiffusion.sample() with sampling_timesteps=steps"""
            gd.sampling_timesteps, gd.is_ddim_sampling = steps, True
            t0 = time.perf_counter()
            with torch.no_grad():
                gd.sample(batch_size=b, init=init[:b], control=control[:b])
            return (time.perf_counter() - t0) / steps
        return "reference", run
    from oracle import diffusion as D
    from oracle.unet3d import Unet3DOracle
    from wdno_b200.unet3d import Unet3D_with_Conv3D
    torch.manual_seed(0)
    orc = Unet3DOracle(Unet3D_with_Conv3D(dim=64, dim_mults=(1, 2, 4), channels=42).state_dict())
    sch = D.schedule("sigmoid", 1000)

    def run(b, steps):
        gg = torch.Generator().manual_seed(7)
        t0 = time.perf_counter()
        with torch.no_grad():
            D.smoke_ddim_sample(orc, sch, (b,) + CONFIGS["C3"]["shape"], steps, 1.0,
                                lambda s: torch.randn(s, generator=gg), [18, 34, 34], init[:b], control[:b])
        return (time.perf_counter() - t0) / steps
    return "port", run


def cpu_rate(cfg_name, steps, warmup, threads, budget_s):
    """DDIM steps/s of the reference's CPU path at the configuration's batch.  `warmup` then `steps` steps are run; if the
    first warm-up step predicts more than `budget_s` for the whole run, the remaining steps use a smaller batch (bounded
    sample) and the rate is scaled to the configuration's batch -- stated in `sample`."""
    import torch
    torch.set_num_threads(threads)
    kind, run = _cpu_models(cfg_name)
    B = CONFIGS[cfg_name]["b"]
    t_first = run(B, 1)                       # warm-up step 1 at the full batch (also sizes the sample)
    total = (max(warmup, 1) - 1 + steps)
    b = B
    if t_first * total > budget_s:
        b = max(1, min(B, int(B * budget_s / (t_first * total))))
    if warmup > 1:
        run(b, warmup - 1)
    per = run(b, steps)
    rate = (b / B) / per
    what = (f"{'reference modules (smoke/ddpm/diffusion_2d.py GaussianDiffusion.sample over video_diffusion_pytorch_conv3d.Unet3D_with_Conv3D)' if kind == 'reference' else 'oracle port of the reference (plain torch fp32)'} "
            f"on {threads} host threads: {max(warmup, 1)} warm-up + {steps} timed DDIM steps at batch {b} "
            f"({per:.2f} s per step; first warm-up step at batch {B}: {t_first:.2f} s)"
            + ("" if b == B else f"; bounded sample: rate scaled by {b}/{B} to the batch-{B} workload"))
    return rate, kind, what, b == B


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg = CONFIGS[args.config]
    K, W = max(1, args.steps), max(1, args.warmup)
    rate, kind, what, full = cpu_rate(args.config, K, W, threads, args.cpu_budget)
    line = {"impl": "reference", "metric": cfg["metric"], "value": rate, "unit": "steps/s", "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["workload"], "same_batch_as_engine_arm": full},
            "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": threads, "kind": kind, "sample": what},
            "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# =============================================================================== engine arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds the reference arm may spend (bounded sample beyond)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None, help="chain length of the e2e leg (default: the configuration's)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, Wm = max(1, args.steps), max(args.warmup, 3)
    cfg = CONFIGS[args.config]

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    from wdno_b200 import _timing, parallel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    m, gd, post, extra = build_engine(args.config, dev)
    B = cfg["b"]
    total = B * world
    conds_h = {k: v.pin_memory() for k, v in host_conditions(args.config, total).items()}
    conds_d = {k: v.to(dev) for k, v in conds_h.items()}

    def run_chain(steps, conds, with_post=True):
        """the product path: steps-long chain on this rank's shard -> fields -> all-gather"""
        gd.sampling_timesteps = steps
        gd.is_ddim_sampling = steps < gd.num_timesteps
        torch.manual_seed(4321)   # same seed on every rank: sharded noise = rows of the single-process draw
        return parallel.sample_sharded(gd, total, post=post if with_post else None, **conds, **extra)

    with torch.no_grad():
        # ---------------- value: W warm-up steps, clock settle, then EXACTLY K timed steps; conditions resident in HBM
        run_chain(Wm, conds_d)              # W warm-up steps: captures the CUDA graphs, builds plans
        t0 = time.time()
        while time.time() - t0 < 1.0:       # let the SM clock settle under load (same chain length as the timed call)
            run_chain(K, conds_d)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            fields = run_chain(K, conds_d)
            e1.record()
            barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        value = world * K / (ms / 1e3)
        launches_per_step = (getattr(gd, "last_launches_per_step", None) or (m.engine().launches + 2)) + 1   # + noise fill
        fields_shape = list(fields.shape)
        del fields

        # ---------------- per-kernel timing: CUDA events around every launch of the dominant kernels, eager steps
        n_prof = 3
        gd.use_cuda_graph = False
        gd.__dict__.pop("_runners", None)
        _timing.sink = None
        run_chain(1, conds_d, with_post=False)       # plan building for the eager runner, untimed
        _timing.sink = []
        run_chain(n_prof, conds_d, with_post=False)
        torch.cuda.synchronize()
        recs = _timing.sink
        _timing.sink = None
        gd.use_cuda_graph = True
        gd.__dict__.pop("_runners", None)
        pk = peaks()
        kern = {}
        for r in recs:
            d = kern.setdefault(r["kernel"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            d["ms"] += r["e0"].elapsed_time(r["e1"]) / n_prof
            d["flops"] += r["flops"] / n_prof
            d["bytes"] += r["bytes"] / n_prof
            d["launches"] += 1
        for d in kern.values():
            d["launches"] //= n_prof
        step_ms = ms / K
        tg = kern.get("tapgemm", dict(ms=1e-9, flops=0.0, launches=0))
        achieved = tg["flops"] / (tg["ms"] * 1e-3) / 1e12
        # DRAM bytes of the same launches: static, from the committed ncu capture (not measured in this run)
        traffic, traffic_src = None, None
        for name in ("r2_tapgemm_traffic.json", "r1d_tapgemm_traffic.json"):
            tp = os.path.join(ROOT, "profiles", name)
            if args.config == "C3" and os.path.exists(tp):
                td = json.load(open(tp))
                if td.get("tapgemm", {}).get("launches") == tg["launches"]:
                    traffic, traffic_src = td["tapgemm"]["dram_bytes_total"], f"static: profiles/{name} (ncu dram__bytes_read+write, same launches)"
                    break
        sub = {}
        for name, d in kern.items():
            if name == "tapgemm":
                continue
            rec = {"launches_per_step": d["launches"], "ms_per_step": d["ms"], "share_of_step": d["ms"] / step_ms}
            if d["bytes"] > 0:
                gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9
                rec.update(bound="hbm", algorithmic_bytes_per_step=d["bytes"], achieved_GBps=gbs, peak_GBps=pk["hbm"],
                           frac=gbs / pk["hbm"])
            if d["flops"] > 0:
                rec["algorithmic_gflop_per_step"] = d["flops"] / 1e9
            sub[name] = rec
        roof = {"bound": "tensor", "kernel": "wdno::tapgemm_kernel (all %d launches of one step)" % tg["launches"],
                "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                "frac_of_burst_peak": achieved / pk["burst"] if pk["burst"] else None, "peak_burst": pk["burst"],
                "peak_source": f"{pk['src']}: cuBLAS bf16 sustained (a kernel timed inside a long step); fp16 operands run at the bf16 rate; "
                               "see `clocks` for the SM clock of this run vs the clock the peak was measured at",
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel_ms_per_step": tg["ms"], "kernel_share_of_step": tg["ms"] / step_ms,
                "algorithmic_gflop_per_step": tg["flops"] / 1e9, "other_kernels": sub}

        # ---------------- DWT sub-records (graph-replayed, HBM roofline; algorithmic bytes: SURVEY.md 8d)
        try:
            roof["dwt"] = dwt_records(dev, pk["hbm"])
        except Exception as e:  # noqa: BLE001 - never lose the headline over a side record
            roof["dwt"] = {"error": f"{type(e).__name__}: {e}"}

        # ---------------- e2e: the configuration's full chain through the public API, HOST conditions in, FIELDS out
        Ke = args.e2e_steps or cfg["chain"]
        out_h = torch.empty(fields_shape, dtype=torch.float32).pin_memory() if rank == 0 else None
        run_chain(Wm, conds_h)              # graphs for the graph runner again (the eager leg dropped them)
        barrier()
        s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        s0.record()
        fields = run_chain(Ke, conds_h)
        s1.record()
        if rank == 0:
            out_h.copy_(fields, non_blocking=True)
        s2.record()
        barrier()
        ems = max_over_ranks(s0.elapsed_time(s2))
        h2d = sum(v.numel() * 4 for v in conds_h.values()) / world     # this rank's shard of every condition
        e2e = {"value": world * Ke / (ems / 1e3), "unit": "steps/s", "steps": Ke,
               "h2d_bytes_per_step": h2d / Ke, "d2h_bytes_per_step": (out_h.numel() * 4 if rank == 0 else 0) / Ke,
               "ms_total": ems, "ms_d2h_of_gathered_fields_rank0": s1.elapsed_time(s2),
               "what": f"parallel.sample_sharded(GaussianDiffusion, batch {total}, post=coefficients->fields) with "
                       f"sampling_timesteps={Ke}: pinned host conditions in (each rank copies its shard), inverse transforms, "
                       f"one all-gather of the fields {fields_shape}, rank 0 copies them to pinned host memory -- all timed"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": cfg["metric"], "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate, f32 state", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "per_gpu_batch": B,
                       "parallelism": f"batch-sharded x{world} (parallel.sample_sharded): no data-path collective, one all-gather of the final fields",
                       "timed_region": f"sample_sharded with sampling_timesteps={K}: initial noise, {K} steps, state->fields, all-gather",
                       "l2": "per-step activation working set (>1 GB at batch 16) exceeds the 126 MB L2; no flush needed",
                       "cuda_graph": True, "rng": "torch Philox stream; sharded ranks evaluate their rows of the full-batch draw"},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches_per_step * K, "roofline": roof,
            "algorithmic_tflops_whole_step": cfg["flops"] * B * world / (step_ms * 1e-3) / 1e12,
            "whole_step_frac_of_sustained_peak": cfg["flops"] * B / (step_ms * 1e-3) / 1e12 / pk["sustained"],
            "whole_step_frac_of_burst_peak": cfg["flops"] * B / (step_ms * 1e-3) / 1e12 / pk["burst"] if pk["burst"] else None}
    if world == 1 and not args.no_cpu_baseline and args.config == "C3":
        threads = os.cpu_count() or 1
        rate, kind, what, _ = cpu_rate("C3", 2, 1, threads, 45.0)
        line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": threads, "kind": kind, "sample": what}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dwt_records(dev, hbm):
    """graph-replayed wavedec3 / waverec3 (80 fields = batch 16 x 5) and Burgers DWTForward / DWTInverse (batch 256)"""
    import torch
    from wdno_b200 import wavelets as W

    def timed_graph(fn, reps=50):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3   # us

    out = {}
    x = torch.randn(80, 32, 64, 64, device=dev)
    coef = W.wavedec3(x, "bior1.3")
    nbytes = x.numel() * 4 + (coef[0].numel() + sum(v.numel() for v in coef[1].values())) * 4
    for name, fn in (("wavedec3", lambda: W.wavedec3(x, "bior1.3")), ("waverec3", lambda: W.waverec3(coef, "bior1.3"))):
        us = timed_graph(fn)
        out[name] = {"us": us, "algorithmic_bytes": nbytes, "achieved_GBps": nbytes / us / 1e3, "peak_GBps": hbm,
                     "frac": nbytes / us / 1e3 / hbm, "workload": "80 fields [32,64,64] (batch 16 x 5), bior1.3 zero"}
    xb = torch.randn(256, 2, 81, 120, device=dev)
    fwd, inv = W.DWTForward(J=1, wave="bior2.4", mode="periodization"), W.DWTInverse(wave="bior2.4", mode="periodization")
    yl, yh = fwd(xb)
    nb = xb.numel() * 4 + (yl.numel() + yh[0].numel()) * 4
    for name, fn in (("DWTForward", lambda: fwd(xb)), ("DWTInverse", lambda: inv((yl, yh)))):
        us = timed_graph(fn)
        out[name] = {"us": us, "algorithmic_bytes": nb, "achieved_GBps": nb / us / 1e3, "peak_GBps": hbm,
                     "frac": nb / us / 1e3 / hbm, "workload": "Burgers batch 256 x 2 fields [81,120], bior2.4 periodization"}
    return out


if __name__ == "__main__":
    main()
