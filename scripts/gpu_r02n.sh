#!/bin/bash
# visit r02n: why the sweep launcher always picks CW = 4; ncu --set full of a late and a mid sweep launch with the source page exported as CSV
set -u
O=gpurun_out/r02n
mkdir -p $O
B="python bench.py --filter ekf --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
SLAM_DEBUG_SWEEP=1 timeout 300 python bench.py --filter ekf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/dbg.json 2> $O/dbg.err
sort $O/dbg.err | uniq -c | sort -rn | head -30 > $O/dbg_summary.txt; cat $O/dbg_summary.txt | head -12
for t in 0 64 128; do python bench.py --filter ekf --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --cta-threads $t 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('cta', $t, 'value %.4g ms %.2f' % (d['value'], d['ms_per_step']))"; done | tee $O/cta_ab.txt
cap() {  # name skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_sweep_kernel -s $2 -c 1 -o /tmp/prof_$1 -f $B > $O/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > $O/$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python scripts/ncu_summary.py full /tmp/prof_$1.ncu-rep $O/$1_full.txt > /dev/null 2>&1
  python scripts/sass_profile.py $O/$1_source.csv live_ekf_slam_b200/csrc/ekf_batch.o ekf_sweep_kernelILi4ELb0 30 > $O/$1_lines.txt 2>&1
  head -12 $O/$1_lines.txt
}
cap late 140
cap mid 112
