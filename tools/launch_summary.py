"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (and optionally list launches)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
show = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
rows = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
    rows.append((name, v, row.get("Grid Size", ""), row.get("Block Size", "")))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("total %.1f us over %d launches" % (tot, len(rows)))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%4d %10.1f us %5.1f%%" % (k[:44], n, v, 100 * v / tot))
for i, (name, v, g, b) in enumerate(rows[:show]):
    print("%4d %-40s %9.1f us grid %s block %s" % (i, name[:40], v, g, b))
