#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_sim_parity.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -12
