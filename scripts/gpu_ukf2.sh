#!/bin/bash
# UKF generation-2 visit: parity tests, quick bench, ncu launch list at late steps; large-map launch list
set -u
O=gpurun_out/r01f
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ukf_parity.py -m gpu -x -q > $O/pytest_ukf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ukf.log
tail -15 $O/pytest_ukf.log
timeout 600 python scripts/quick_bench.py 4096 1000 ukf 2 2>&1 | tail -3 | tee $O/quick_ukf.txt
U="python bench.py --filter ukf --steps 1 --warmup 3 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 500 --csv --log-file $O/launches_ukf.csv $U > $O/ncu_launch_ukf.log 2>&1
python scripts/ncu_summary.py launch $O/launches_ukf.csv $O/launches_ukf.txt; cat $O/launches_ukf.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 600 --csv --log-file $O/launches_large.csv python scripts/bench_large.py 2000 3000 300 > $O/ncu_launch_large.log 2>&1
python scripts/ncu_summary.py launch $O/launches_large.csv $O/launches_large.txt; cat $O/launches_large.txt
