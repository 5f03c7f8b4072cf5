"""Aggregate an ncu gpu__time_duration CSV by kernel name (template arguments kept): count, total ms, share.
Usage: kernel_shares.py launches.csv [N]   (N: only the last N launches, e.g. one training step)"""
import csv
import sys
from collections import defaultdict

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
tot = defaultdict(lambda: [0, 0.0])
rows = list(csv.DictReader(lines))
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
for r in rows:
    name = r["Kernel Name"]
    short = name[5 if name.startswith("void ") else 0:name.find("(")]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    tot[short][0] += 1
    tot[short][1] += us
s = sum(v[1] for v in tot.values())
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {n:5d} launches {us/1000:9.2f} ms {100*us/s:5.1f} %")
print(f"total {s/1000:.2f} ms")
