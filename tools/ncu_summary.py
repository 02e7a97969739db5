"""key metrics + stall samples of ncu reports: python tools/ncu_summary.py rep1.ncu-rep [...] (build container)"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def summary(rep):
    o = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(o.splitlines()))
    hdr, unit = rd[0], rd[1]
    out = []
    for val in rd[2:]:
        d = dict(zip(hdr, val))
        u = dict(zip(hdr, unit))
        lines = [f"### `{d.get('Kernel Name', '')[:110]}`", "", "| metric | unit | value |", "|---|---|---:|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {u[k]} | {d[k]} |")
        st = []
        for k, v in d.items():
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    st.append((float(v), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(a for a, _ in st) or 1.0
        lines.append("")
        lines.append("warp-state samples: " + ", ".join(f"{n} {100 * a / tot:.0f}%" for a, n in sorted(st, reverse=True)[:8]))
        out.append("\n".join(lines))
    return "\n\n".join(out)


if __name__ == "__main__":
    for r in sys.argv[1:]:
        print(summary(r))
        print()
