#!/bin/bash
# Round-1 closing visit: smoke(), the whole GPU suite, and the bench lines of the final kernels.
set -u
O=gpurun_out/r01h
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_ekf.json 2> $O/bench_ekf.err; echo "ekf rc=$?"
timeout 900 python bench.py --filter ukf --steps 1 --warmup 3 > $O/bench_ukf.json 2> $O/bench_ukf.err; echo "ukf rc=$?"
timeout 900 python scripts/bench_large.py 2000 3000 300 2> $O/bench_large.err | grep "^{" > $O/bench_large.json; echo "large rc=$?"
wc -l $O/bench_ekf.json $O/bench_ukf.json $O/bench_large.json
