"""Markdown table of an ncu launch list (--metrics gpu__time_duration.sum --csv):
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("sy2d::", "").replace("(int)", "").replace("void ", "")
    t = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[iu], 1e-6)
    agg[name][0] += 1
    agg[name][1] += t


def table(items, title, total):
    print(f"| kernel | launches | total ms | {title} |\n|---|---:|---:|---:|")
    for k, (n, ms) in sorted(items, key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")


tot = sum(v[1] for v in agg.values())
print(f"{len(rows) - 1} launches, {tot:.2f} ms under ncu\n")
table(agg.items(), "share", tot)
grid = [(k, v) for k, v in agg.items() if "k_problem" not in k]
print("\nWithout the ensemble kernel (the single-grid part of the bench):\n")
table(grid, "share of the grid part", sum(v[1] for _, v in grid))
