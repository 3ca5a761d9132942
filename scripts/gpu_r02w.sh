#!/bin/bash
# visit r02w: ncu --set full of the new sweep kernel: one-warp CTAs on a small tile, four-warp CTAs on the full tile (source pages as CSV)
set -u
O=gpurun_out/r02w
mkdir -p $O
B="python bench.py --filter ekf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
cap() {  # name regex skip symbol
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -o /tmp/prof_$1 -f $B > $O/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > $O/$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python scripts/ncu_summary.py full /tmp/prof_$1.ncu-rep $O/$1_full.txt > /dev/null 2>&1
  INNER=1 python scripts/sass_profile.py $O/$1_source.csv live_ekf_slam_b200/csrc/ekf_batch.o $4 40 > $O/$1_lines.txt 2>&1
  head -8 $O/$1_lines.txt
}
cap cw1 "ekf_sweep_kernel<.int.1, .bool.0>" 8 ekf_sweep_kernelILi1ELb0
export SLAM_TUNE="0=50"
cap cw4 "ekf_sweep_kernel<.int.4, .bool.0>" 39 ekf_sweep_kernelILi4ELb0
tail -n 2 $O/ncu_cw1.log; tail -n 2 $O/ncu_cw4.log
