#!/bin/bash
# ncu captures of the two EKF kernels at steady state (late in a 1000-step sweep)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# sweep kernel: 4th launch = the timed one (3 warm-ups before)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_sweep_kernel -s 3 -c 1 -o gpurun_out/prof_sweep -f $B > gpurun_out/ncu_sweep.log 2>&1
# step kernel late in a per-step sweep (2 matching launches per filter step: main + retry)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_step_kernel -s 1900 -c 4 -o gpurun_out/prof_step -f $B --no-sweep > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file gpurun_out/launches_step.csv $B --no-sweep > gpurun_out/ncu_launch_step.log 2>&1
tail -2 gpurun_out/ncu_sweep.log gpurun_out/ncu_step.log
