#!/bin/bash
# ncu launch list (per-kernel durations, serialised) of bin/mTR on N C5 reads in ONE group.  Usage: bash tools/gpu_ncu_list.sh N tag
set -u
N=${1:-256}; TAG=${2:-r2}
mkdir -p gpurun_out
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads($N, seed=1000)
synth.write_fasta('/tmp/c5n.fa', reads, line_width=0)
PY
t0=$(date +%s.%N)
MTR_GROUPS_PER_GPU=1 MTR_GROUP_READS=$N bin/mTR -c /tmp/c5n.fa > /tmp/c5n.out 2> gpurun_out/${TAG}_plain.err
t1=$(date +%s.%N)
python -c "print('plain run wall s', $t1-$t0)"
grep -E "all$|Computing|wrap around|count table" gpurun_out/${TAG}_plain.err
MTR_GROUPS_PER_GPU=1 MTR_GROUP_READS=$N timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${TAG}_launches.csv bin/mTR /tmp/c5n.fa > /tmp/c5n_ncu.out 2>/dev/null
python tools/summarise_launches.py gpurun_out/${TAG}_launches.csv "bin/mTR on $N C5 reads, one group" > gpurun_out/${TAG}_launches_summary.csv
cat gpurun_out/${TAG}_launches_summary.csv | head -50
rm -f gpurun_out/${TAG}_launches.csv
