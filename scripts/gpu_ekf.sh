#!/bin/bash
# EKF GPU visit: parity tests of the batched EKF path + a quick sweep timing
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_sim_parity.py tests/test_gpu_cpp_host.py tests/test_gpu_large_map.py -m gpu -x -q > gpurun_out/pytest_ekf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ekf.log
tail -30 gpurun_out/pytest_ekf.log
timeout 300 python scripts/quick_bench.py 4096 1000 ekf 3 2>&1 | tail -4
