#!/usr/bin/env python
"""Condense ncu artefacts from gpurun_out/ into small, committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r01_launches.md
    python profiles/summarize.py full gpurun_out/prof.ncu-rep profiles/r01_ncu_mono8.md [--traffic-voxels N]
"""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r[ki], [0, 0.0, []])
        v = float(r[vi].replace(",", ""))
        a[0] += 1
        a[1] += v
        a[2].append(v)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({os.path.basename(src)}): gpu__time_duration.sum, --clock-control none\n\n")
        f.write("Cold-cache, serialised per-launch times: compare SHARES, not absolutes.\n\n")
        f.write("| launches | total ms | share | max ms | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {max(a[2]) / 1e6:.3f} | `{k[:110]}` |\n")
    print(open(dst).read())


def full(src, dst, voxels=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(src)})\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            f.write("\n")
        d = dict(zip(hdr, rows[-1]))
        if voxels:
            def tobytes(key):
                v, u = float(d[key].replace(",", "")), units[hdr.index(key)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            tr = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
            json.dump({"source": os.path.basename(dst), "voxels": voxels, "dram_bytes_per_launch": tr,
                       "dram_bytes_per_voxel": tr / voxels}, open(os.path.join(os.path.dirname(dst), "traffic.json"), "w"))
            f.write(f"DRAM traffic per launch: {tr / 1e9:.3f} GB = {tr / voxels:.2f} B/voxel over {voxels} voxels "
                    f"(algorithmic: 44 B/voxel)\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        vox = int(sys.argv[sys.argv.index("--traffic-voxels") + 1]) if "--traffic-voxels" in sys.argv else None
        full(sys.argv[2], sys.argv[3], vox)
