#!/usr/bin/env python
"""Latency probe for concurrent DP lanes: thread A runs long batches back to back on one context while thread B times
tiny batches on another context (separate streams, same GPU).  Prints wall time per tiny batch next to the kernel's
CUDA-event time, with the long lane idle / running fused / running split."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mtr_b200 import capi, synth
import wdp_cases

rng = np.random.default_rng(5)
reads = [synth.long_reads(1, seed=50 + i)[0][0] for i in range(8)]
tails = [(0, 0)] * len(reads)
packed, woff, lens = capi.pack_reads(reads, tails)

def jobs_for(n, lo, hi, umax):
    js = []
    for _ in range(n):
        r = int(rng.integers(len(reads))); L = len(reads[r])
        rows = min(int(rng.integers(lo, hi + 1)), L - 2); first = int(rng.integers(0, L - rows))
        ulen = int(rng.integers(2, min(umax, rows // 5 + 2) + 1))
        js.append(dict(read=r, first=first, rows=rows, unit=rng.integers(0, 4, ulen).astype(np.uint8), gain=1, mis=1, indel=3))
    return wdp_cases.build_job_array(js, pair=True)

tiny = jobs_for(300, 20, 128, 25)
long_ = jobs_for(450, 3000, 16000, 450)
if os.environ.get('DRY'): sys.exit(0)
a, b = capi.Context(0), capi.Context(0)
a.upload_reads(packed, woff, lens)
b.lib.mtr_reads_share(b.h, a.h); b.n_reads = a.n_reads
stop = False
def hog(fused):
    a.wdp_set_fused_traceback(fused)
    n = 0; t0 = time.perf_counter()
    while not stop:
        a.wdp_run(long_[0], long_[1]); n += 1
    st = a.stats()
    print("   long lane: %d batches, %.2f ms/batch wall, kernel fill %.2f tb %.2f ms" % (n, (time.perf_counter() - t0) / max(n, 1) * 1e3, st["wdp_fill_ms"], st["wdp_tb_ms"]), flush=True)

def probe(label):
    for _ in range(20): b.wdp_run(tiny[0], tiny[1])
    ts, ks = [], []
    for _ in range(200):
        t0 = time.perf_counter(); b.wdp_run(tiny[0], tiny[1]); ts.append(time.perf_counter() - t0)
        st = b.stats(); ks.append(st["wdp_fill_ms"] + st["wdp_tb_ms"])
    ts = np.array(ts) * 1e3
    print("%-34s tiny batch wall ms: p50 %.3f p90 %.3f max %.3f | kernel ms p50 %.3f" % (label, np.median(ts), np.percentile(ts, 90), ts.max(), np.median(ks)), flush=True)

for tiny_fused in (True, False):
    b.wdp_set_fused_traceback(tiny_fused)
    probe("tiny fused=%s, long lane idle" % tiny_fused)
    for fused in (True, False):
        stop = False
        th = threading.Thread(target=hog, args=(fused,)); th.start(); time.sleep(0.2)
        probe("tiny fused=%s, long lane fused=%s" % (tiny_fused, fused))
        stop = True; th.join()


# ---- the directional index of 1024 reads looping on a third context next to the DP lanes
print("--- DP lane latency next to a running directional index", flush=True)
many = [synth.long_reads(256, seed=90)[0][i % 256] for i in range(1024)]
pk, wo, ln = capi.pack_reads(many, [(0, 0)] * len(many))
d = capi.Context(0)
d.upload_reads(pk, wo, ln)
def di_hog():
    n = 0; t0 = time.perf_counter()
    while not stop:
        d.di_run(True); n += 1
    print("   directional index: %d runs of 1024 reads, %.1f ms each" % (n, (time.perf_counter() - t0) / max(n, 1) * 1e3), flush=True)
b.wdp_set_fused_traceback(True)
for prio in (0, 3):
    b.lib.mtr_set_priority(b.h, prio)
    stop = False
    th = threading.Thread(target=di_hog); th.start(); time.sleep(0.3)
    probe("tiny batch, DI running, DP stream priority %d" % prio)
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < 0.5:
        b.wdp_run(long_[0], long_[1]); k += 1
    print("   long batch next to DI: %.2f ms each" % ((time.perf_counter() - t0) / k * 1e3), flush=True)
    stop = True; th.join()
