#!/bin/bash
# quick GPU visit: parity tests + one bench line per filter
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ekf.json 2> gpurun_out/bench_ekf.err; tail -3 gpurun_out/bench_ekf.err
cat gpurun_out/bench_ekf.json
