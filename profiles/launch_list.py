"""Render an `ncu --metrics gpu__time_duration.sum --csv` launch list as a table: per kernel (name, grid) the number of
launches, total and mean duration.  Usage: python profiles/launch_list.py gpurun_out/r02_launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {k: i for i, k in enumerate(hdr)}
acc = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("agx::", "").replace("(int)", "").replace("(bool)", "")
    grid = r[ix["Grid Size"]].strip("()").split(",")[0].strip()
    ns = float(r[ix["Metric Value"]].replace(",", ""))
    if r[ix["Metric Unit"]] in ("us", "usecond"):
        ns *= 1e3
    a = acc.setdefault((name, grid), [0, 0.0])
    a[0] += 1
    a[1] += ns
print("| kernel | grid | launches | total us | mean us |\n|---|---|---|---|---|")
for (name, grid), (n, ns) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:90]}` | {grid} | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.1f} |")
