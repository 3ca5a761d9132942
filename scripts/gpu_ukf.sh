#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ukf_parity.py -m gpu -x -q > gpurun_out/pytest_ukf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ukf.log
tail -30 gpurun_out/pytest_ukf.log
timeout 600 python scripts/quick_bench.py 4096 1000 ukf 2 2>&1 | tail -4
