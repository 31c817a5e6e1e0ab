#!/usr/bin/env python
"""Worker CPU of the host pipeline per C5 read, measured WITHOUT a GPU: the product's pipeline.cpp linked against the
test-only simulated device (tests/hostsim) runs N synthetic C5 reads with MTR_PROFILE=1 and the profile's stage timers
(k-mer count tables, max-node listing, de Bruijn walks, polish, everything else) are printed per read.  The simulated
device answers DP calls with the CPU oracle on the dispatcher threads, so only the `host cpu-s` line is meaningful.

    python tools/host_cpu_probe.py [reads=96] [worker threads=2] [repeats=3]
"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtr_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
threads = sys.argv[2] if len(sys.argv) > 2 else "2"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
simdir = os.path.join(ROOT, "tests", "hostsim")
subprocess.check_call(["make", "-s", "-C", simdir])
reads = synth.long_reads(n, seed=1003)[0]
with tempfile.NamedTemporaryFile(suffix=".fa") as f:
    synth.write_fasta(f.name, reads, ids=list(range(n)))
    best = None
    for _ in range(reps):
        p = subprocess.run([os.path.join(simdir, "_build", "mTR_hostsim"), f.name], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                           env=dict(os.environ, MTR_PROFILE="1", MTR_THREADS=threads))
        m = re.search(r"host cpu-s: step ([\d.]+) build ([\d.]+) maxlist ([\d.]+) walk ([\d.]+) polish ([\d.]+) \| chains (\d+) walks (\d+)", p.stderr.decode())
        v = [float(x) for x in m.groups()[:5]] + [int(m.group(6)), int(m.group(7))]
        if best is None or v[0] < best[0]:
            best = v
step, build, lst, walk, pol, chains, walks = best
print("%d reads, best of %d: %.2f ms worker CPU per read = tables %.2f + listing %.2f + walks %.2f + polish %.2f + rest %.2f   (%d chains, %d walks per read)"
      % (n, reps, step / n * 1e3, build / n * 1e3, lst / n * 1e3, walk / n * 1e3, pol / n * 1e3, (step - build - lst - walk - pol) / n * 1e3, chains // n, walks // n))
