#!/bin/bash
set -u
O=gpurun_out/r01f
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_large_map.py tests/test_gpu_sim_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python scripts/bench_large.py 2000 3000 300 2>&1 | tail -1 | tee $O/bench_large.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 600 --csv --log-file $O/launches_large.csv python scripts/bench_large.py 2000 3000 300 > $O/ncu_launch_large.log 2>&1
python scripts/ncu_summary.py launch $O/launches_large.csv $O/launches_large.txt; cat $O/launches_large.txt
