"""Aggregate an `ncu --page source --csv` dump: stall-reason totals, their distribution over the code, and shared
memory excess wavefronts.  Usage: python profiles/srcstat.py dump.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_")]
data = [r for r in rows[h + 1:] if len(r) >= len(hdr) and r[0] != "Address"]


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


samples = sum(num(r[idx["# Samples"]]) for r in data)
print("kernel:", rows[0][1] if rows[0] else "?", "| instructions:", len(data), "| samples:", samples)
tot = {s: sum(num(r[idx[s]]) for r in data) for s in stalls}
for s, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v:
        print(f"  {s:24s} {v:8d} {100.0 * v / max(samples, 1):5.1f}%")
nb = 12
print("samples / no_inst / long_sb / barrier per code segment:")
for b in range(nb):
    seg = data[b * len(data) // nb:(b + 1) * len(data) // nb]
    print(f"  seg{b:02d}", sum(num(r[idx['# Samples']]) for r in seg), sum(num(r[idx['stall_no_inst']]) for r in seg),
          sum(num(r[idx['stall_long_sb']]) for r in seg), sum(num(r[idx['stall_barrier']]) for r in seg))
exc = sum(num(r[idx["L1 Wavefronts Shared Excessive"]]) for r in data)
wf = sum(num(r[idx["L1 Wavefronts Shared"]]) for r in data)
print("shared wavefronts:", wf, "excessive:", exc)
