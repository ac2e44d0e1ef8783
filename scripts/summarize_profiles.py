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


def build(src, dst, title, peak_gbs="6553.3"):
    """per-kernel DRAM traffic / achieved bandwidth of ONE build (a csv of `ncu --metrics duration,dram bytes` over
    scripts/trace_build.py): launches between the last leaf_init and the end."""
    peak = float(peak_gbs)
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row["ID"], row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v * 1e6 if unit == "s" else v  # -> us
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        per.setdefault(key, {})[row["Metric Name"]] = v
    rows = [(k[1], m) for k, m in per.items()]
    idx = [i for i, (n, _) in enumerate(rows) if "leaf_init" in n or "presplit_aabb" in n]
    rows = rows[idx[-1]:] if idx else rows
    agg = collections.OrderedDict()
    for n, m in rows:
        n = re.sub(r"\(.*", "", n)
        n = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::", "", n)
        a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    tot_t = sum(a[1] for a in agg.values())
    tot_b = sum(a[2] + a[3] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n")
        f.write("Source: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` over one "
                "`build_cwbvh_from_tris` (third build of scripts/trace_build.py: arena and result cache warm).\n")
        f.write("Per-launch times under ncu are serialised and cold-cache; achieved GB/s = DRAM bytes / that time, against the measured "
                f"copy bandwidth of {peak:.0f} GB/s (MEASURED_PEAKS.json).\n\n")
        f.write(f"Total: {len(rows)} launches, {tot_t / 1000:.2f} ms of kernel time, {tot_b / 1e9:.2f} GB of DRAM traffic "
                f"({tot_b / 1e9 / (tot_t * 1e-6) if tot_t else 0:.0f} GB/s average, {100 * tot_b / 1e9 / (tot_t * 1e-6) / peak if tot_t else 0:.0f} % of peak).\n\n")
        f.write("| kernel | launches | time us | share | DRAM read MB | DRAM write MB | GB/s | % of HBM peak |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for n, (c, t, r, w) in sorted(agg.items(), key=lambda x: -x[1][1]):
            gbs = (r + w) / 1e9 / (t * 1e-6) if t else 0.0
            f.write(f"| `{n[:90]}` | {c} | {t:.1f} | {100 * t / tot_t:.1f}% | {r / 1e6:.1f} | {w / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f}% |\n")
    print(dst, len(rows), "launches", f"{tot_t:.1f} us")


if __name__ == "__main__":
    {"launches": launches, "full": full, "build": build}[sys.argv[1]](*sys.argv[2:6])
