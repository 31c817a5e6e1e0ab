#!/usr/bin/env python
"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total ms, share."""
import csv, re, sys
from collections import defaultdict
rows = defaultdict(lambda: [0, 0.0])
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
head = next(r)
ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
for row in r:
    if len(row) <= vi:
        continue
    name = re.sub(r"\(.*", "", row[ki])
    v = float(row[vi].replace(",", ""))
    unit = row[ui]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    rows[name][0] += 1
    rows[name][1] += ms
tot = sum(v[1] for v in rows.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes")
print("kernel,launches,total_ms,share")
for k, (n, ms) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.3f,%.4f" % (k, n, ms, ms / tot))
