#!/bin/bash
# ncu full capture of ekf_sweep_kernel launches: late chunks (full tile) and early chunks (small tile)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_sweep_kernel -s 52 -c 3 -o gpurun_out/prof_sweep_late -f $B > gpurun_out/ncu_sweep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_sweep_kernel -s 10 -c 2 -o gpurun_out/prof_sweep_early -f $B >> gpurun_out/ncu_sweep.log 2>&1
tail -3 gpurun_out/ncu_sweep.log
