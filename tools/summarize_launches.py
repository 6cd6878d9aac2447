"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: mean microseconds per kernel."""
import collections
import csv
import sys

lines = [ln for ln in open(sys.argv[1]) if not ln.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg.setdefault(name, []).append(v)
tot = 0.0
for k, v in agg.items():
    if k.startswith("void at::"):
        continue
    m = sum(v[1:]) / max(1, len(v) - 1) if len(v) > 1 else v[0]
    tot += m
    print(f"{k[:48]:48s} n={len(v):3d} mean={m:9.1f} us")
print(f"{'sum of means (one step)':48s}       {tot:9.1f} us")
