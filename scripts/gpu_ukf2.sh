#!/bin/bash
set -u
O=gpurun_out/r01g
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ukf_parity.py tests/test_gpu_loc_naive.py -m gpu -x -q > $O/pytest_ukf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ukf.log
tail -8 $O/pytest_ukf.log
timeout 600 python scripts/quick_bench.py 4096 1000 ukf 1 2 1 2>&1 | tail -1 | tee $O/quick_ukf.txt
U="python bench.py --filter ukf --steps 1 --warmup 3 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 7600 -c 160 --csv --log-file $O/launches_ukf.csv $U > $O/ncu_launch_ukf.log 2>&1
python scripts/ncu_summary.py launch $O/launches_ukf.csv $O/launches_ukf.txt; cat $O/launches_ukf.txt
