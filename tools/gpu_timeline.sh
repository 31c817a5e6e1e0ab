#!/bin/bash
# per-stage latency of the engine's waves on N C5 reads: MTR_TIMELINE stamps summarised by tools/timeline.py
N=${1:-16384}; TAG=${2:-tl}; shift 2
mkdir -p gpurun_out
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads($N, seed=1000)
synth.write_fasta('/tmp/c5n.fa', reads, line_width=0)
PY
env MTR_TIMELINE=1 MTR_PROFILE=1 MTR_GROUP_MBASES=100000 "$@" timeout 600 bin/mTR -c /tmp/c5n.fa 2> gpurun_out/${TAG}.err | md5sum
grep -v "wave [0-9]*:\|^\[timeline\]" gpurun_out/${TAG}.err | cut -c1-900 | tail -16
python tools/timeline.py gpurun_out/${TAG}.err
grep " wave [0-9]*:" gpurun_out/${TAG}.err | sed -E 's/.*wave ([0-9]+): unfinished ([0-9]+) accepted ([0-9]+) tasks short ([0-9]+) \(slots ([0-9+]+)\) long ([0-9]+) \(slots ([0-9+]+)\) deferred ([0-9]+) in flight ([0-9]+)/\1:u\2:s\4:l\6(\7):d\8:f\9/' | awk 'NR<=40 || NR%6==0' | tr '\n' ' '
