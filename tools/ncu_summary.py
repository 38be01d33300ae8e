"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
    python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep profiles/x_r1.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        a = agg.setdefault(re.sub(r"\(.*", "", r[ki])[:110], [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: {src}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over ONE step "
                f"of bench.py (cold-cache, serialised: compare shares, not absolutes).  {len(rows) - 1} launches, "
                f"{tot:.2f} ms total.\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {v:.3f} | {100 * v / tot:.1f}% | {n} | `{k}` |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {src}\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')][:120]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for i, h in enumerate(hdr):
                if h in KEEP:
                    f.write(f"| {h} | {r[i]} | {units[i]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
