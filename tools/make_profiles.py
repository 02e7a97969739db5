"""Turn gpurun_out/ artefacts into the tracked summaries under profiles/ (run in the build container)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launch_summary(src, dst_csv, dst_md, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    per = []
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        per.append((re.sub(r"\(.*", "", r["Kernel Name"]), v))
    idx = [i for i, (n, _) in enumerate(per) if "step_begin" in n]
    step = per[idx[-1]:] if idx else per
    with open(dst_csv, "w") as f:
        f.write("launch,kernel,ms\n")
        for i, (n, v) in enumerate(step):
            f.write(f"{i},{n},{v:.6f}\n")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for n, v in step:
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    with open(dst_md, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over the last CUDA-graph-replayed DDIM step "
                f"(cold-cache, serialised per-launch times: compare SHARES).\n\n{len(step)} launches, {T:.3f} ms summed.\n\n"
                "| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{k[:80]}` | {cnt[k]} | {v:.3f} | {100 * v / T:.1f}% |\n")


def ncu_raw(rep, dst, title):
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    o = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(o.splitlines()))
    hdr, unit, val = rd[0], rd[1], rd[2]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nfrom `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`, one launch)\n\n"
                "| metric | unit | value |\n|---|---|---:|\n")
        for h, u, v in zip(hdr, unit, val):
            if h in keys or h in ("Kernel Name",):
                f.write(f"| {h} | {u} | {v} |\n")


def traffic_summary(src, dst):
    """per-launch DRAM bytes of the tap-GEMM / conv1x1 launches of one step -> json (bench.py reads the totals)"""
    import json
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]))
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
            per.setdefault(k, {})["us"] = v
        elif r["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            per.setdefault(k, {})["read" if "read" in r["Metric Name"] else "write"] = v * mult
    out = {"what": "DRAM traffic of the tap-GEMM and conv1x1 launches of one Unet3D forward (C3, batch 16): ncu --metrics "
                   "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none (tools/gpu_prof_r1d.sh)"}
    for tag in ("tapgemm", "conv1x1"):
        rows = [(k, v) for k, v in per.items() if tag in k[1]]
        out[tag] = {"launches": len(rows), "dram_bytes_read": sum(v.get("read", 0) for _, v in rows),
                    "dram_bytes_write": sum(v.get("write", 0) for _, v in rows),
                    "dram_bytes_total": sum(v.get("read", 0) + v.get("write", 0) for _, v in rows),
                    "sum_duration_us_cold": sum(v.get("us", 0) for _, v in rows),
                    "per_launch": [{"id": int(k[0]), "us": round(v.get("us", 0), 2), "read_MB": round(v.get("read", 0) / 1e6, 2),
                                    "write_MB": round(v.get("write", 0) / 1e6, 2)} for k, v in rows]}
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    t = os.path.join(GO, f"tapgemm_traffic_{TAG}.csv")
    if os.path.exists(t):
        traffic_summary(t, os.path.join(OUT, f"{TAG}_tapgemm_traffic.json"))
    s = os.path.join(GO, f"launches_{TAG}_step.csv")
    if os.path.exists(s):
        launch_summary(s, os.path.join(OUT, f"{TAG}_launches_step.csv"), os.path.join(OUT, f"{TAG}_launches_step.md"),
                       f"{TAG}: per-kernel split of one DDIM step (C3, batch 16)")
    for name, title in (("prof_tapgemm_c64", "tap-GEMM, 3x3x3 64->64 conv, B=16 24x40x40 (dominant kernel)"),
                        ("prof_conv1x1", "conv1x1 kernel, 128->64 res_conv at B=16 24x40x40 (HBM-bound)"),
                        ("prof_tapgemm_stem_strips", "tap-GEMM, 7x7x7 82(96)->64 stem of the super-resolution model in "
                                                     "column-strip mode, B=4 24x80x80")):
        rep = os.path.join(GO, f"{name}_{TAG}.ncu-rep")
        if os.path.exists(rep):
            ncu_raw(rep, os.path.join(OUT, f"{TAG}_{name}.md"), title)
    for extra in ("tapgemm_breakdown_C3.json", "tapgemm_breakdown_C2.json", "tapgemm_breakdown_C4.json", "bench_configs.jsonl",
                  "step_profile_C3.txt", "step_profile_C2.txt", "step_profile_C4.txt"):
        p = os.path.join(GO, extra)
        if os.path.exists(p):
            open(os.path.join(OUT, f"{TAG}_{extra}"), "w").write(open(p).read())
