#!/bin/bash
# Round-2 visit B: the restructured bench line (top level + configs.ukf / .large / .mixed) and the touched parity tests.
set -u
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_ukf_parity.py -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 > $O/bench_all.json 2> $O/bench_all.err ) 2> $O/bench_all.time; echo "bench rc=$?"
tail -3 $O/bench_all.time; tail -5 $O/bench_all.err
wc -c $O/bench_all.json
