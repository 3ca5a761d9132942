#!/bin/bash
# ncu --set full captures (with source counters) of the UKF back kernel and the tridiagonalisation kernel late in the sweep
set -u
O=gpurun_out/r02j
mkdir -p $O
U="python bench.py --filter ukf --steps 1 --warmup 0 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ukf_back3_kernel -s 1800 -c 1 -o $O/prof_back3 -f $U > $O/ncu_back3.log 2>&1
tail -2 $O/ncu_back3.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ukf_front2_kernel -s 2702 -c 1 -o $O/prof_front2 -f $U > $O/ncu_front2.log 2>&1
tail -2 $O/ncu_front2.log
