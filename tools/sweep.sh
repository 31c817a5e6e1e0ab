#!/bin/bash
# tuning aid: bench.py --quick over a list of "ENV=... ENV=..." settings (one per line on stdin or in $1)
# usage: bash tools/sweep.sh tag steps warmup < settings.txt
TAG=${1:-sweep}; STEPS=${2:-2}; WARM=${3:-1}
mkdir -p gpurun_out
: > gpurun_out/${TAG}.log
while IFS= read -r line; do
  [ -z "$line" ] && continue
  echo "## $line" >> gpurun_out/${TAG}.log
  env $line timeout 150 python bench.py --quick --steps $STEPS --warmup $WARM >> gpurun_out/${TAG}.log 2>> gpurun_out/${TAG}.err
  echo "rc=$?" >> gpurun_out/${TAG}.log
done
cat gpurun_out/${TAG}.log
