"""Print the launches of an ncu gpu__time_duration CSV in order: kernel (template arguments), grid size, microseconds."""
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for i, r in enumerate(csv.DictReader(lines)):
    name = r["Kernel Name"]
    short = name[name.find("void ") + 5 if "void " in name else 0:name.find("(")]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    print(f"{i:4d} {short:45s} grid {r['Grid Size']:>14s} {us:9.1f} us")
