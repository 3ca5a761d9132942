#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_st -s 1900 -c 2 -o gpurun_out/prof_step -f $B --no-sweep > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 4000 --csv --log-file gpurun_out/launches_step.csv $B --no-sweep > gpurun_out/ncu_launch_step.log 2>&1
tail -n 2 gpurun_out/ncu_step.log
