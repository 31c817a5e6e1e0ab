#!/bin/bash
# Round-2 ncu evidence (run on the GPU box; results land in gpurun_out/, the summaries are copied to profiles/ by hand):
#   1. launch list of `python bench.py --quick` (every kernel of the timed region, serialised: compare SHARES)
#   2. DRAM traffic of every K3 launch of the same command (metrics-only pass) -> bytes per cell run
#   3. ncu --set full of K3 launches (first wave = full GPU, and mid-run) and of the directional-index kernels
set -u
R=${1:-2048}; TAG=${2:-r2}
mkdir -p gpurun_out
BENCH="python bench.py --quick --steps 1 --warmup 1 --reads $R"
$BENCH > gpurun_out/${TAG}_quick_plain.json 2>/dev/null; cat gpurun_out/${TAG}_quick_plain.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_quick_ncu1.json 2>/dev/null
python tools/summarise_launches.py gpurun_out/${TAG}_launches.csv "$BENCH (warm-up step + timed step)" > gpurun_out/${TAG}_launches_summary.csv
head -40 gpurun_out/${TAG}_launches_summary.csv
gzip -f gpurun_out/${TAG}_launches.csv
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:wdp_fill_family --clock-control none -c 60000 --csv --log-file gpurun_out/${TAG}_k3_dram.csv $BENCH > gpurun_out/${TAG}_quick_ncu2.json 2>/dev/null
python - <<PY
import csv, json
rows = [l for l in open('gpurun_out/${TAG}_k3_dram.csv') if l.startswith('"')]
r = list(csv.reader(rows)); h = r[0]
ki, mi, vi, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
tot = {}; n = 0
for x in r[1:]:
    v = float(x[vi].replace(',', ''))
    u = x[ui]
    if x[mi].startswith('dram'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    else:
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1}.get(u, 1)
        n += 1
    tot[x[mi]] = tot.get(x[mi], 0) + v
q = json.loads(open('gpurun_out/${TAG}_quick_ncu2.json').read().strip().splitlines()[-1])
cells = q['cells_run'] * 2          # warm-up step + timed step run the same number of reads (different seeds: approximately)
out = {'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over all %d wdp_fill_family launches of `$BENCH` (tools/gpu_profile_r2.sh)' % n,
       'dram_bytes_read': tot.get('dram__bytes_read.sum'), 'dram_bytes_write': tot.get('dram__bytes_write.sum'), 'k3_launches': n,
       'k3_ms_serialised': tot.get('gpu__time_duration.sum'), 'cells_run_approx': cells,
       'dram_bytes_per_cell': (tot.get('dram__bytes_read.sum', 0) + tot.get('dram__bytes_write.sum', 0)) / max(cells, 1),
       'algorithmic_dir_bytes_per_cell': q['dir_bytes'] / max(q['cells_run'], 1)}
json.dump(out, open('gpurun_out/${TAG}_k3_traffic.json', 'w'), indent=1)
print(json.dumps(out))
PY
rm -f gpurun_out/${TAG}_k3_dram.csv
python - <<PY
import sys
sys.path.insert(0, '.')
from mtr_b200 import synth
reads, _ = synth.long_reads(8192, seed=1000)
synth.write_fasta('/tmp/c5n.fa', reads, line_width=0)
PY
M=gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__icc_request_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed
for st in no_instruction wait barrier long_scoreboard short_scoreboard math_pipe_throttle not_selected branch_resolving dispatch_stall lg_throttle mio_throttle selected membar; do M=$M,smsp__average_warps_issue_stalled_${st}_per_issue_active.ratio; done
cap() {   # name, kernel regex, skip, count
  MTR_GROUPS_PER_GPU=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$2" -s $3 -c $4 -o gpurun_out/${TAG}_$1 -f bin/mTR /tmp/c5n.fa > /dev/null 2> gpurun_out/${TAG}_$1.err
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv --metrics $M 2>/dev/null > gpurun_out/${TAG}_$1_raw.csv
  python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/${TAG}_$1_raw.csv')))
hdr = rows[0]
skip = ('ID','Process ID','Process Name','Host Name','Context','Stream','Block Size','Grid Size','Device','CC','Section Name','Metric Name','Metric Unit')
keys = [k for k in hdr if k not in skip and k != 'Kernel Name']
print('| metric | ' + ' | '.join((dict(zip(hdr, r)).get('Kernel Name', '')[:34]).replace('|', '/') for r in rows[2:]) + ' |')
print('|---|' + '---|' * len(rows[2:]))
for k in keys:
    print('| %s | ' % k + ' | '.join(dict(zip(hdr, r)).get(k, '') for r in rows[2:]) + ' |')
PY
}
cap k3_first "wdp_fill_family" 0 4 > gpurun_out/${TAG}_k3_first_summary.md
cap k3_mid "wdp_fill_family" 400 4 > gpurun_out/${TAG}_k3_mid_summary.md
cap di "di_" 0 12 > gpurun_out/${TAG}_di_summary.md
cap eng "eng_sched|eng_unitfinder" 40 6 > gpurun_out/${TAG}_eng_summary.md
head -30 gpurun_out/${TAG}_k3_first_summary.md | cut -c1-250
