#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small summaries committed under profiles/.

    python tools/ncu_summaries.py launches gpurun_out/launches_r01d.csv r01d "bench.py --steps 2 --warmup 3 --no-cpu-baseline"
        -> profiles/<tag>_ncu_launch_summary.csv, profiles/<tag>_traffic.json
    python tools/ncu_summaries.py full gpurun_out/prof_r01d.ncu-rep r01d "note"
        -> profiles/<tag>_ncu_full_summary.csv   (needs `ncu` on PATH to read the report)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "launch__grid_size", "launch__block_size",
]


def launches(path, tag, cmd):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        if len(r) > vi:
            per[(r[ii], r[ki])][r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (_, k), m in per.items():
        a = agg[k.split("(")[0]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_summary.csv")
    with open(out, "w") as f:
        f.write(f"# {tag} ncu launch list summary (`{cmd}`, B200)\n")
        f.write("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes; "
                "dram_MB = dram__bytes_read+write per launch\n")
        f.write("kernel,launches,avg_us,share,dram_MB_per_launch\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k},{a[0]},{a[1] / a[0] / 1e3:.1f},{a[1] / tot:.4f},{a[2] / a[0] / 1e6:.1f}\n")
    import re
    m = re.search(r"--config (c\d)", cmd)
    config = m.group(1) if m else "c2"
    batch = {"c1": 64, "c2": 64, "c3": 32, "c4": 16}[config]            # bench.py CONFIGS
    tj = {"source": f"profiles/{tag}_ncu_launch_summary.csv (ncu launch list of `{cmd}`, batch {batch} {config.upper()} frames)",
          "frames_per_launch": batch, "config": config,
          "dram_bytes_per_launch": {k.replace("cm3d::", ""): int(a[2] / a[0]) for k, a in agg.items() if k.startswith("cm3d::")}}
    with open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w") as f:
        json.dump(tj, f, indent=1)
    print(open(out).read())


def full(path, tag, note):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in FULL_METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv")
    with open(out, "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none --import-source on; {note}\n")
        f.write("Kernel Name," + ",".join(hdr[c] for c in cols) + "\n")
        f.write("," + ",".join(units[c] for c in cols) + "\n")
        for r in rows[2:]:
            f.write(r[ki].split("(")[0] + "," + ",".join(r[c].replace(",", "") for c in cols) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
