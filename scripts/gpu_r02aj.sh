#!/bin/bash
# ncu --set full of the four-warp sweep kernel at the full tile after the storage-order rank-2 walk
set -u
O=gpurun_out/r02aj; mkdir -p $O
export SLAM_TUNE="0=50"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ekf_sweep_kernel<.int.4, .bool.0>" -s 39 -c 1 -o /tmp/prof_sweep -f python bench.py --filter ekf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_sweep.log 2>&1
unset SLAM_TUNE
python scripts/ncu_summary.py full /tmp/prof_sweep.ncu-rep $O/ekf_sweep_full.txt > /dev/null 2>&1
ncu -i /tmp/prof_sweep.ncu-rep --page source --csv > $O/sweep_source.csv 2>/dev/null
ncu -i /tmp/prof_sweep.ncu-rep --page raw --csv > $O/sweep_raw.csv 2>/dev/null
INNER=1 python scripts/sass_profile.py $O/sweep_source.csv live_ekf_slam_b200/csrc/ekf_batch.o ekf_sweep_kernelILi4ELb0 50 > $O/ekf_sweep_lines.txt 2>&1
head -14 $O/ekf_sweep_lines.txt
timeout 120 python scripts/sweep_chunks.py 0 | head -1 | cut -c1-60
