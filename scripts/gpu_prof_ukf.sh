#!/bin/bash
# UKF: bench line at the full 1000 steps + ncu full capture of late-step ukf_step_kernel launches
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --filter ukf --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ukf_full.json 2> gpurun_out/bench_ukf_full.err; tail -3 gpurun_out/bench_ukf_full.err
cat gpurun_out/bench_ukf_full.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ukf_step_kernel -s 1150 -c 2 -o gpurun_out/prof_ukf -f \
   python bench.py --filter ukf --steps 1 --warmup 3 --filter-steps 300 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ukf.log 2>&1
tail -2 gpurun_out/ncu_ukf.log
