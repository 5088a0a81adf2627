"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: per-kernel share of device time.
Usage: python tools/summarize_launches.py launches.csv "<command that was profiled>" > summary.txt"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0].replace("void ", "")[:70]
    unit = r.get("Metric Unit", "ns")
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("command:", sys.argv[2] if len(sys.argv) > 2 else "?")
print("per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
print(f"{'total us':>12s} {'share':>7s} {'launches':>8s}  kernel")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:12.1f} {100 * t / tot:6.2f}% {c:8d}  {n}")
