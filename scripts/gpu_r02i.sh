#!/bin/bash
set -u
O=gpurun_out/r02i
mkdir -p $O
U="python bench.py --filter ukf --steps 1 --warmup 0 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ukf_eig3_kernel' -s 950 -c 1 -o $O/prof_eig3 -f $U > $O/ncu.log 2>&1
tail -2 $O/ncu.log
