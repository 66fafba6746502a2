"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, average and total time.
usage: python tools/launch_summary.py launches.csv "<the command that was profiled>" """
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += t
print(f"ncu launch list of `{sys.argv[2] if len(sys.argv) > 2 else '?'}` (cold caches, serialised):")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:4d} x {k[:76]:76s} {t / c:9.1f} us avg {t:10.1f} us total")
