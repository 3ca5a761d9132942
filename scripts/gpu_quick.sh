#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_sim_parity.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['accuracy'])"
