#!/bin/bash
# Round-2 development aid: GPU tests + a first timing of the resident engine on C5 reads.  Usage (on the GPU box):
#   bash tools/gpu_check.sh [n_reads] [pytest -k expression]
set -u
N=${1:-2048}
K=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/gpu.txt 2>&1
if [ "$K" = "none" ]; then
  echo "pytest skipped" > gpurun_out/pytest_gpu.log
elif [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads($N, seed=1000)
synth.write_fasta('/tmp/c5.fa', reads, line_width=0)
synth.write_fasta('/tmp/c5_64.fa', reads[:64], line_width=0)
print('bases', sum(len(r) for r in reads))
PY
for g in ${GROUPS_LIST:-1024}; do
  for k in ${CTX_LIST:-8}; do
    t0=$(date +%s.%N)
    env MTR_PROFILE=${PROFILE:-1} MTR_GROUP_READS=$g MTR_GROUPS_PER_GPU=$k timeout 600 bin/mTR -c /tmp/c5.fa > /tmp/c5.out 2> gpurun_out/c5_g${g}_k${k}.err
    rc=$?
    t1=$(date +%s.%N)
    echo "group_reads=$g contexts=$k rc=$rc md5=$(md5sum < /tmp/c5.out) wall=$(echo "$t1 - $t0" | bc -l 2>/dev/null || python -c "print($t1-$t0)") s"
    grep -E "all$|ranges|Computing|wrap around|count table|chaining|Count of" gpurun_out/c5_g${g}_k${k}.err
  done
done
timeout 600 bin/mTR /tmp/c5_64.fa | md5sum
timeout 600 oracle/_ref/mTR_ref_det /tmp/c5_64.fa | md5sum
