#!/usr/bin/env python
"""Diagnostic: bin/mTR on a FASTA file of B batches of R C5 reads with MTR_PROFILE=1; prints wall time and the per-batch
profile lines (which engine phases overlap, where the time goes)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
texts = bench.make_texts(R, 0, B)
path = "/dev/shm/e2e_probe.fa"
with open(path, "wb") as f:
    for t in texts:
        f.write(t)
env = dict(os.environ, MTR_PROFILE="1", MTR_BATCH_READS=str(R), MTR_BATCH_MBASES=str(max(64, (R * 21000) >> 20)))
t0 = time.perf_counter()
p = subprocess.run([os.path.join(ROOT, "bin", "mTR"), path], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
dt = time.perf_counter() - t0
print("mTR: %d reads in %.2f s (%.0f reads/s incl. start-up), rc %d" % (R * B, dt, R * B / dt, p.returncode))
for line in p.stderr.decode().splitlines():
    if "finish times" in line or "host cpu" in line or "batch " in line:
        print(line[:220])
os.unlink(path)
