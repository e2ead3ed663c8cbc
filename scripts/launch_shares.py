#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total ms, share. Usage: launch_shares.py <csv> [skip]"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("ccb::", "")
    ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[iu], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
print("%d launches, %.3f ms in total (cold-cache, serialised ncu replay: shares, not absolute times)" % (len(rows) - 1 - skip, tot / 1e6))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%%  %5d x %10.3f ms  %s" % (100 * ns / tot, n, ns / n / 1e6, name[:150]))
