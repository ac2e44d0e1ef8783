#!/usr/bin/env python
"""Turns gpurun_out ncu artefacts into small tracked summaries under profiles/.

  python scripts/summarize_profiles.py launches gpurun_out/launches.csv profiles/r1_launches_kitchen.md "title"
  python scripts/summarize_profiles.py full gpurun_out/prof_traverse.ncu-rep profiles/r1_traverse_full.md "title"
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v * 1e6 if unit == "s" else v
        rows.append((row["Kernel Name"], v))
    # last complete step: between the last two leaf_init launches (one build_cwbvh_from_tris each)
    idx = [i for i, (n, _) in enumerate(rows) if "leaf_init" in n]
    # the timed steps come before the e2e steps; take the 4th build (first timed step after 3 warm-ups) when present
    start, end = (idx[3], idx[4]) if len(idx) > 4 else (idx[-2], idx[-1]) if len(idx) > 1 else (0, len(rows))
    step = rows[start:end]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        n = re.sub(r"\(.*", "", n)
        n = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::", "", n)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v for _, v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n")
        f.write(f"Source: `ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py ...` ({len(rows)} launches captured).\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
        f.write(f"One step (build_cwbvh_from_tris + traversal + bench bookkeeping) = launches {start}..{end}: {len(step)} launches, {tot:.1f} us of kernel time.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{n[:100]}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |\n")
    print(dst, len(step), "launches", f"{tot:.1f} us")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` ({src}), read with `ncu -i ... --page raw --csv`.\n\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name[:120]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for h, u, v in zip(hdr, units, vals):
                if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                    f.write(f"| {h} | {v} | {u} |\n")
            f.write("\n")
    print(dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
