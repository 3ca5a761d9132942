#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full capture of the top kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ekf.json 2> gpurun_out/bench_ekf.err
timeout 600 python bench.py --filter ukf --steps 2 --warmup 3 --filter-steps 200 --no-cpu-baseline > gpurun_out/bench_ukf.json 2> gpurun_out/bench_ukf.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_ekf.csv \
   python bench.py --steps 1 --warmup 3 --filter-steps 300 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_step_kernel -s 900 -c 3 -o gpurun_out/prof_ekf -f \
   python bench.py --steps 1 --warmup 3 --filter-steps 300 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ekf.json
