#!/usr/bin/env python
"""Summarises MTR_TIMELINE output ([timeline] ctx wave t0..t10, %globaltimer ns): per-stage latency of a wave."""
import sys, collections
import numpy as np
names = ["begin", "advance", "polish", "sched", "walk", "emit", "plan", "scatter", "zero_aux", "-", "publish"]
rows = collections.defaultdict(list)
for line in open(sys.argv[1]):
    if not line.startswith("[timeline]"):
        continue
    f = line.split()
    try:
        rows[f[1]].append([int(x) for x in f[2:14]])
    except ValueError:
        continue
seq = [0, 1, 2, 3, 5, 6, 7, 8, 10]
print("groups", len(rows), "waves", sum(len(v) for v in rows.values()))
tot = collections.defaultdict(list)
wave_len = []
for ctx, v in rows.items():
    a = np.array(v, dtype=np.int64)
    a = a[a[:, 1] > 0]
    for i in range(len(seq) - 1):
        d = (a[:, 1 + seq[i + 1]] - a[:, 1 + seq[i]]) / 1e6
        tot["%s->%s" % (names[seq[i]], names[seq[i + 1]])] += list(d[(a[:, 1 + seq[i]] > 0) & (a[:, 1 + seq[i + 1]] > 0)])
    nxt = (a[1:, 1] - a[:-1, 1 + 10]) / 1e6
    tot["publish->next begin"] += list(nxt[(a[1:, 1] > 0) & (a[:-1, 11] > 0)])
    wave_len += list((a[1:, 1] - a[:-1, 1]) / 1e6)
    w = (a[:, 1 + 4] - a[:, 1 + 3]) / 1e6
    tot["sched->walk start"] += list(w[(a[:, 5] > 0) & (a[:, 4] > 0)])
for k, d in tot.items():
    d = np.array(d)
    if len(d):
        print("%-24s mean %8.3f ms  median %8.3f  p90 %8.3f  max %8.3f  sum %9.1f" % (k, d.mean(), np.median(d), np.percentile(d, 90), d.max(), d.sum()))
wl = np.array(wave_len)
print("wave period              mean %8.3f ms  median %8.3f  p90 %8.3f" % (wl.mean(), np.median(wl), np.percentile(wl, 90)))
