#!/bin/bash
# tuning aid: runs bench.py --quick under several settings, one per line on stdin: "VAR=.. VAR=.. [-- bench args]"
out=${1:-gpurun_out/tune.log}
: > $out
while read -r line; do
  envs="${line%%--*}"; args=""
  case "$line" in *--*) args="${line#*--}";; esac
  echo "### $line" >> $out
  env $envs MTR_PROFILE=1 python bench.py --quick --warmup 1 --steps 4 $args >> $out 2> gpurun_out/tune_last.err
  grep "tier\|finish\|host cpu" gpurun_out/tune_last.err | grep -v "tier 6" | tail -5 >> $out
  grep "timeline" gpurun_out/tune_last.err | tail -16 >> $out
done
