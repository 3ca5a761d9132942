#!/bin/bash
set -u
O=gpurun_out/r01f
mkdir -p $O
U="python bench.py --filter ukf --steps 1 --warmup 3 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ukf_back2_kernel<13>|ukf_back2_kernel' -s 1900 -c 1 -o $O/prof_back2 -f $U > $O/ncu_back2.log 2>&1
tail -2 $O/ncu_back2.log
