#!/usr/bin/env python
"""Turns one gpu_round.sh visit (gpurun_out/<tag>_*) into the tracked evidence under profiles/:
  <tag>_launches.csv        the ncu launch list as captured (gpu__time_duration.sum per launch)
  <tag>_launch_shares.txt   per-kernel count / mean ns / share of the listed time
  <tag>_kernel_metrics.txt  selected raw metrics of the integrate kernel from the --set full capture
  <tag>_hot_lines.txt       per-source-line shares of instructions and stall samples
  <tag>_bench.json          the bench line of the same visit (NOT taken under a profiler)
usage: python tools/ncu_summary.py <tag>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
is_bench_workload = "--other" not in sys.argv  # pass --other for captures of workloads that are not bench.py's
frames_per_launch = int(sys.argv[sys.argv.index("--frames-per-launch") + 1]) if "--frames-per-launch" in sys.argv else 16
G, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def num(s):
    return float(s.replace(",", ""))


for stem in ("launches", "bench_launches"):
    src = os.path.join(G, f"{tag}_{stem}.csv")
    if not os.path.exists(src):
        continue
    shutil.copy(src, os.path.join(PR, f"{tag}_{stem}.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[iK], []).append(num(r[iV]))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(PR, f"{tag}_{stem.replace('launches', 'launch_shares')}.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        f.write(f"{'kernel':70} {'launches':>8} {'mean_ns':>10} {'share':>7}\n")
        for k, v in agg.items():
            f.write(f"{k[:70]:70} {len(v):8d} {sum(v) / len(v):10.0f} {sum(v) / tot:7.3f}\n")

rep = os.path.join(G, f"{tag}_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PR, f"{tag}_kernel_metrics.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; one column per captured launch\n")
        iN = hdr.index("Kernel Name")
        f.write(f"kernel: {rows[2][iN]}\n")
        traffic = []
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k:90} {units[i]:12} " + "  ".join(r[i] for r in rows[2:]) + "\n")
        iR, iW = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        for r in rows[2:]:
            traffic.append(num(r[iR]) * scale[units[iR]] + num(r[iW]) * scale[units[iW]])
        iT = hdr.index("gpu__time_duration.sum")
    cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::1"],
                        capture_output=True, text=True).stdout
    tmp = os.path.join("/tmp", f"{tag}_cs.csv")
    open(tmp, "w").write(cs)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp, "0.6"], capture_output=True, text=True).stdout
    open(os.path.join(PR, f"{tag}_hot_lines.txt"), "w").write("# per source line: share of executed warp instructions, share of stall samples, top stall reasons\n" + hot)

b = os.path.join(G, f"{tag}_bench.json")
if os.path.exists(b) and os.path.getsize(b):
    shutil.copy(b, os.path.join(PR, f"{tag}_bench.json"))
print(sorted(os.listdir(PR)))
