#!/bin/bash
# ncu --set full of K3 launches (first wave = full GPU, and mid-run), of the directional-index kernels and of the engine's
# scheduler / unit-finder kernels; only the metric tables come back (the .ncu-rep files stay on the box)
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads(8192, seed=1000)
synth.write_fasta('/tmp/c5n.fa', reads, line_width=0)
PY
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__icc_request_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__maximum_warps_per_active_cycle_pct
for st in no_instruction wait barrier long_scoreboard short_scoreboard math_pipe_throttle not_selected branch_resolving dispatch_stall lg_throttle mio_throttle selected membar; do M=$M,smsp__average_warps_issue_stalled_${st}_per_issue_active.ratio; done
cap() {   # name, kernel regex, skip, count
  MTR_GROUPS_PER_GPU=1 timeout 600 ncu --set full --clock-control none -k regex:"$2" -s $3 -c $4 -o /tmp/${TAG}_$1 -f bin/mTR /tmp/c5n.fa > /dev/null 2> gpurun_out/${TAG}_$1.err
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv --metrics $M 2>/dev/null > gpurun_out/${TAG}_$1_raw.csv
  python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/${TAG}_$1_raw.csv')))
hdr = rows[0]
skip = ('ID','Process ID','Process Name','Host Name','Context','Stream','Block Size','Grid Size','Device','CC','Section Name','Metric Name','Metric Unit')
keys = [k for k in hdr if k not in skip and k != 'Kernel Name']
print('ncu --set full --clock-control none -k regex:$2 -s $3 -c $4 bin/mTR <8192 C5 reads, one context> (tools/gpu_profile_r2b.sh); units: ' + ', '.join('%s [%s]' % (k, u) for k, u in zip(hdr, rows[1]) if k in ('gpu__time_duration.sum', 'dram__bytes_read.sum')))
print()
print('| metric | ' + ' | '.join((dict(zip(hdr, r)).get('Kernel Name', '')[:34]).replace('|', '/') for r in rows[2:]) + ' |')
print('|---|' + '---|' * len(rows[2:]))
for k in keys:
    print('| %s | ' % k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ' (warps per issue)') + ' | '.join(dict(zip(hdr, r)).get(k, '') for r in rows[2:]) + ' |')
PY
  rm -f /tmp/${TAG}_$1.ncu-rep
}
cap k3_first "wdp_fill_family" 0 4 > gpurun_out/${TAG}_k3_first_summary.md
cap k3_mid "wdp_fill_family" 300 4 > gpurun_out/${TAG}_k3_mid_summary.md
cap di "di_" 0 12 > gpurun_out/${TAG}_di_summary.md
cap eng "eng_sched|eng_unitfinder" 60 6 > gpurun_out/${TAG}_eng_summary.md
wc -c gpurun_out/${TAG}_*_summary.md
