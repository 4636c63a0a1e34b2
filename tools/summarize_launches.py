#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"].split("(")[0], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
agg = collections.OrderedDict()
for name, ns, grid, blk in rows:
    a = agg.setdefault(name, [0, 0.0, grid, blk]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}  grid block")
for name, (n, ns, grid, blk) in agg.items():
    print(f"{name[:60]:60s} {n:5d} {ns/1e3:10.1f} {ns/1e3/n:9.1f} {100*ns/tot:5.1f}%  {grid} {blk}")
# last evaluation + SPR batch in launch order
print("\nlast 14 launches:")
for name, ns, grid, blk in rows[-14:]:
    print(f"  {name[:60]:60s} {ns/1e3:9.1f} us  {grid} {blk}")
