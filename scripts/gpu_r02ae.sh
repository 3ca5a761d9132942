#!/bin/bash
# ncu --set full of the two-warp sweep kernel on a mid-size tile
set -u
O=gpurun_out/r02ae
mkdir -p $O
B="python bench.py --filter ekf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
cap() {  # name regex skip symbol
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o /tmp/prof_$1 -f $B > $O/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > $O/$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python scripts/ncu_summary.py full /tmp/prof_$1.ncu-rep $O/$1_full.txt > /dev/null 2>&1
  INNER=1 python scripts/sass_profile.py $O/$1_source.csv live_ekf_slam_b200/csrc/ekf_batch.o $4 40 > $O/$1_lines.txt 2>&1
  python scripts/sass_profile.py $O/$1_source.csv live_ekf_slam_b200/csrc/ekf_batch.o $4 16 > $O/$1_outer.txt 2>&1
  head -8 $O/$1_lines.txt
}
cap cw2 "ekf_sweep_kernel<.int.2, .bool.0>" 12 ekf_sweep_kernelILi2ELb0
