#!/bin/bash
# Round-2 visit A: the whole GPU suite with the new parity tests (sigma points, parameter variants, full-size large map).
set -u
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -40 $O/pytest_gpu.log
