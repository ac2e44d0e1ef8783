"""python scripts/kernel_times.py <ncu csv> -- per-kernel time / DRAM totals of the LAST build in an ncu csv of scripts/trace_build.py"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.OrderedDict()
for r in csv.DictReader(rows):
    k = (r["ID"], r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    per.setdefault(k, {})[r["Metric Name"]] = v
items = [(k[1], m) for k, m in per.items()]
idx = [i for i, (n, _) in enumerate(items) if "leaf_init" in n]
items = items[idx[-1]:] if idx else items
agg = collections.OrderedDict()
for n, m in items:
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::", "", n)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us, {len(items)} launches")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
    print(f"  {n[:60]:60s} {c:3d} {t:9.1f} us")
