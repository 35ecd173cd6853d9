"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the
long launches in order.  usage: summarize_launches.py file.csv [min_us]"""
import collections
import csv
import sys

path = sys.argv[1]
min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.OrderedDict()
seq = []
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0]
    t = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit in ("ns", "nsecond"):
        t /= 1e3
    elif unit in ("ms", "msecond"):
        t *= 1e3
    seq.append((name, t, row.get("Grid Size", "")))
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += t
T = sum(v[1] for v in tot.values())
print(f"total {T:.0f} us over {len(seq)} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {100*v[1]/T:5.1f}% x{v[0]:3d}  {k[:90]}")
print()
for i, (n, t, g) in enumerate(seq):
    if t > min_us:
        print(i, f"{t:8.1f}", g, n[:60])
