#!/bin/bash
# Round-1 (visit d) GPU-box visit: parity tests, bench lines for configs[1..3], reference arm, ncu launch lists and
# full captures of the dominant kernels.  Every command is bounded by its own timeout.
set -u
O=gpurun_out/r01d
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.csv
nproc > $O/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_ekf.json 2> $O/bench_ekf.err; echo "ekf rc=$?"
timeout 900 python bench.py --filter ukf --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_ukf.json 2> $O/bench_ukf.err; echo "ukf rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python scripts/bench_large.py 2000 3000 300 > $O/bench_large.json 2> $O/bench_large.err; echo "large rc=$?"
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
# launch list of the default bench command (value path), all launches
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_ekf.csv $B > $O/ncu_launch_ekf.log 2>&1
# full captures: late sweep chunks, late step-kernel launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_sweep_kernel -s 52 -c 2 -o $O/prof_sweep -f $B > $O/ncu_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_step_kernel -s 1900 -c 2 -o $O/prof_step -f $B --no-sweep > $O/ncu_step.log 2>&1
# UKF: launch list (steps 900-1000 of the first sweep) and full capture of late launches of each of the three kernels
U="python bench.py --filter ukf --steps 1 --warmup 3 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 500 --csv --log-file $O/launches_ukf.csv $U > $O/ncu_launch_ukf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ukf_ -s 2850 -c 6 -o $O/prof_ukf -f $U > $O/ncu_ukf.log 2>&1
tail -2 $O/ncu_sweep.log $O/ncu_step.log $O/ncu_ukf.log
cat $O/bench_ekf.json $O/bench_ukf.json $O/bench_large.json
