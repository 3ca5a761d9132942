#!/bin/bash
# ncu full capture (with source) of the three generation-2 UKF kernels at a late step (n ~ 100)
set -u
O=gpurun_out/r01g
mkdir -p $O
U="python bench.py --filter ukf --steps 1 --warmup 3 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ukf_(back2|front2|ql)_kernel' -s 2850 -c 3 -o $O/prof_ukf2 -f $U > $O/ncu_ukf2.log 2>&1
tail -2 $O/ncu_ukf2.log
ls -la $O
