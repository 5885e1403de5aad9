#!/usr/bin/env python
"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, avg us, share."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
c = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0][-52:]
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    t = c.setdefault(n, [0, 0.0])
    t[0] += 1
    t[1] += us
tot = sum(v[1] for v in c.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for n, v in sorted(c.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f | %.3f |" % (n, v[0], v[1], v[1] / v[0], v[1] / tot))
print("\ntotal %.1f us over %d launches" % (tot, sum(v[0] for v in c.values())))
