#!/bin/bash
# where does a tiny input spend its time?  (the GPU tests run the CLI on inputs of a few reads)
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
small, _ = synth.rand_seq_reads(15, 12, 0.02, 0.05, 0.04, 120, 90, 4, seed=5)
synth.write_fasta('/tmp/tiny.fa', small)
PY
for env in "X=1" "MTR_GROUPS_PER_GPU=1" "MTR_PROFILE=1" "X=2"; do
  echo "## $env"
  ( time env $env bin/mTR /tmp/tiny.fa > /dev/null ) 2>&1 | grep -v "wave [0-9]*:" | cut -c1-300 | tail -12
done
