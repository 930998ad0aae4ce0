"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, time, share."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    k = r[ki][:90]
    agg[k][0] += 1
    agg[k][1] += v * scale
tot = sum(v for _, v in agg.values())
print("%-92s %5s %12s %7s" % ("kernel", "n", "total_ms", "share"))
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-92s %5d %12.3f %7.3f" % (k, c, v, v / tot))
