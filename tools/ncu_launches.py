"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean
device time per kernel, share of the listed launches.  ncu_launches.py CSV [first [count]]"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[first:first + count]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("davo::", "")
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + float(r[14]) / 1e3)
tot = sum(t for _, t in agg.values())
print("launches %d..%d of %s (cold-cache, serialised under ncu: compare shares, not absolutes)" % (first, first + len(rows) - 1, sys.argv[1]))
print("%-52s %8s %12s %10s %7s" % ("kernel", "launches", "total us", "mean us", "share"))
for k, (n, t) in agg.items():
    print("%-52s %8d %12.1f %10.2f %6.1f%%" % (k[:52], n, t, t / n, 100 * t / tot))
print("%-52s %8d %12.1f" % ("total", len(rows), tot))
