#!/bin/bash
# one group of N C5 reads with the engine's profile counters (MTR_PROFILE), burst 1 so that every wave is printed
N=${1:-256}
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads($N, seed=1000)
synth.write_fasta('/tmp/c5n.fa', reads, line_width=0)
PY
MTR_PROFILE=1 MTR_ENGINE_BURST=1 MTR_GROUPS_PER_GPU=1 MTR_GROUP_READS=$N bin/mTR -c /tmp/c5n.fa 2> gpurun_out/prof1.err | md5sum
grep -v "wave" gpurun_out/prof1.err | tail -14
grep "wave" gpurun_out/prof1.err | awk '{print $5, $7, $9, $12, $17}' | head -70 | tr '\n' ';'
