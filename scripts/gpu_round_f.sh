#!/bin/bash
# Round-1 (visit f, final) GPU-box visit: smoke(), parity tests, bench lines (EKF, UKF, mixed, reference arm, large map),
# ncu launch lists and full captures of the kernels changed since visit e.  Every command has its own timeout.
set -u
O=gpurun_out/r01f
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.csv
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_ekf.json 2> $O/bench_ekf.err; echo "ekf rc=$?"
timeout 900 python bench.py --filter ukf --steps 1 --warmup 3 > $O/bench_ukf.json 2> $O/bench_ukf.err; echo "ukf rc=$?"
timeout 600 python bench.py --filter mixed --instances 1024 --filter-steps 500 --steps 2 --warmup 3 > $O/bench_mixed.json 2> $O/bench_mixed.err; echo "mixed rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --impl reference --filter ukf --steps 1 --warmup 1 > $O/bench_ref_ukf.json 2> $O/bench_ref_ukf.err; echo "ref ukf rc=$?"
timeout 900 python scripts/bench_large.py 2000 3000 300 > $O/bench_large.json 2> $O/bench_large.err; echo "large rc=$?"
U="python bench.py --filter ukf --steps 1 --warmup 3 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 7600 -c 160 --csv --log-file $O/launches_ukf.csv $U > $O/ncu_launch_ukf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ukf_(back2|front2|ql)_kernel' -s 3800 -c 4 -o $O/prof_ukf2 -f $U > $O/ncu_ukf2.log 2>&1
LG="python scripts/bench_large.py 2000 3000 300"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 600 --csv --log-file $O/launches_large.csv $LG > $O/ncu_launch_large.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lm_(front|gemm)' -s 5600 -c 2 -o $O/prof_large -f $LG > $O/ncu_large.log 2>&1
tail -n 2 $O/ncu_ukf2.log; tail -n 2 $O/ncu_large.log
cat $O/smoke.log | tail -4
