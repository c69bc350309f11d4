"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    k = row["Kernel Name"][:90]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("share    launches  avg_us     kernel   (total %.1f us)" % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%%  %6d  %9.1f  %s" % (v[1] / tot * 100, v[0], v[1] / v[0], k))
