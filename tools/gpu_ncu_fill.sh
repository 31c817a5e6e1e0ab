#!/bin/bash
# ncu --set full of a few K3 launches (fill family kernels + traceback) from the middle of a one-group run
N=${1:-512}; TAG=${2:-r2}; SKIP=${3:-30}
mkdir -p gpurun_out
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads($N, seed=1000)
synth.write_fasta('/tmp/c5n.fa', reads, line_width=0)
PY
MTR_GROUPS_PER_GPU=1 MTR_GROUP_READS=$N timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"wdp_fill_family|wdp_traceback_dev" -s $SKIP -c 6 -o gpurun_out/${TAG}_k3 -f bin/mTR /tmp/c5n.fa > /dev/null 2> gpurun_out/${TAG}_ncu.err
tail -3 gpurun_out/${TAG}_ncu.err
ncu -i gpurun_out/${TAG}_k3.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_no_instruction.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed 2>/dev/null > gpurun_out/${TAG}_k3_raw.csv
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/${TAG}_k3_raw.csv')))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get('Kernel Name','')[:40])
    for k, v in d.items():
        if k in ('ID','Process ID','Process Name','Host Name','Kernel Name','Context','Stream','Block Size','Grid Size','Device','CC','Section Name','Metric Name','Metric Unit'): continue
        print('   %-90s %s' % (k, v))
PY
