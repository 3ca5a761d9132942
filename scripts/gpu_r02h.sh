#!/bin/bash
# ncu full capture of the multi-warp UKF back kernel at a late step
set -u
O=gpurun_out/r02h
mkdir -p $O
U="python bench.py --filter ukf --steps 1 --warmup 0 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ukf_back3_kernel' -s 1900 -c 1 -o $O/prof_back3 -f $U > $O/ncu.log 2>&1
tail -2 $O/ncu.log
