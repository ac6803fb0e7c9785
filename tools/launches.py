#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST `n` launches."""
import csv, re, sys, collections
f = sys.argv[1]; tail = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(f) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rows = list(csv.DictReader(lines))[-tail:]
agg = collections.OrderedDict()
for r in rows:
    k = re.sub(r"\(.*", "", r["Kernel Name"])[:64]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
for k, (c, t, m) in agg.items():
    print(f"{k:64s} x{c:3d} total {t:9.1f} us  max {m:9.1f} us  {100*t/tot:5.1f}%")
print(f"total {tot:.1f} us")
