"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py launches.csv "title" > profiles/xyz.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki][:100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("%-102s %6s %12s %10s %7s" % ("kernel", "count", "total_ms", "avg_ms", "share"))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-102s %6d %12.3f %10.4f %6.1f%%" % (k, c, t, t / c, 100 * t / tot))
