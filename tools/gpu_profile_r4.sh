#!/bin/bash
# Round-2 final ncu evidence, lean (run on the GPU box; results land in gpurun_out/, summaries are copied to profiles/ by hand):
#   1. launch list of `python bench.py --quick --steps 1 --warmup 1 --reads R` (every kernel, serialised: compare SHARES)
#   2. ncu --set full of the first K3 launches of a group (full GPU) with the stall table
set -u
R=${1:-2048}; TAG=${2:-r4}
mkdir -p gpurun_out
BENCH="python bench.py --quick --steps 1 --warmup 1 --reads $R"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_quick_ncu1.json 2>/dev/null
python tools/summarise_launches.py gpurun_out/${TAG}_launches.csv "$BENCH (warm-up step + timed step)" > gpurun_out/${TAG}_launches_summary.csv
head -24 gpurun_out/${TAG}_launches_summary.csv
rm -f gpurun_out/${TAG}_launches.csv
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads(4096, seed=1000)
synth.write_fasta('/tmp/c5n.fa', reads, line_width=0)
PY
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__occupancy_limit_registers,sm__maximum_warps_per_active_cycle_pct
for st in no_instruction wait barrier long_scoreboard short_scoreboard math_pipe_throttle not_selected branch_resolving dispatch_stall selected; do M=$M,smsp__average_warps_issue_stalled_${st}_per_issue_active.ratio; done
MTR_GROUPS_PER_GPU=1 timeout 300 ncu --set full --clock-control none -k regex:wdp_fill_family -s 0 -c 4 -o /tmp/${TAG}_k3 -f bin/mTR /tmp/c5n.fa > /dev/null 2> gpurun_out/${TAG}_k3.err
ncu -i /tmp/${TAG}_k3.ncu-rep --page raw --csv --metrics $M 2>/dev/null > gpurun_out/${TAG}_k3_raw.csv
python - <<PY > gpurun_out/${TAG}_k3_summary.md
import csv
rows = list(csv.reader(open('gpurun_out/${TAG}_k3_raw.csv')))
hdr = rows[0]
skip = ('ID','Process ID','Process Name','Host Name','Context','Stream','Block Size','Grid Size','Device','CC','Section Name','Metric Name','Metric Unit')
keys = [k for k in hdr if k not in skip and k != 'Kernel Name']
print('ncu --set full --clock-control none -k regex:wdp_fill_family -s 0 -c 4 bin/mTR <4096 C5 reads, one context> (tools/gpu_profile_r4.sh); units: ' + ', '.join('%s [%s]' % (k, u) for k, u in zip(hdr, rows[1]) if k in ('gpu__time_duration.sum', 'dram__bytes_read.sum')))
print()
print('| metric | ' + ' | '.join((dict(zip(hdr, r)).get('Kernel Name', '')[:34]).replace('|', '/') for r in rows[2:]) + ' |')
print('|---|' + '---|' * len(rows[2:]))
for k in keys:
    print('| %s | ' % k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ' (warps per issue)') + ' | '.join(dict(zip(hdr, r)).get(k, '') for r in rows[2:]) + ' |')
PY
head -40 gpurun_out/${TAG}_k3_summary.md | cut -c1-220
