"""Summarise the ncu CSV of scripts/layer_probe.py: median duration of the probed layer per variant."""
import csv
import sys

import numpy as np

path, nvar, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 3
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    name = r["Kernel Name"]
    if "conv_umma" in name and ", 1, 0>" not in name:        # skip the fp32 "out" conv of the harness
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((name[name.find("<"):name.find(">") + 1], us))
assert len(rows) == nvar * reps, (len(rows), nvar, reps)
for i in range(nvar):
    chunk = rows[i * reps:(i + 1) * reps]
    print(f"variant {i}: kernel {chunk[0][0]}  median {np.median([c[1] for c in chunk]):8.1f} us   all {[round(c[1], 1) for c in chunk]}")
