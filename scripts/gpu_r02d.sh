#!/bin/bash
# Round-2 visit D: ncu full captures of the three generation-3 UKF kernels at a late step (n ~ 100) + launch list + bench
set -u
O=gpurun_out/r02d
mkdir -p $O
U="python bench.py --filter ukf --steps 1 --warmup 0 --filter-steps 1000 --no-e2e --no-cpu-baseline"
timeout 900 python bench.py --filter ukf --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_ukf.json 2> $O/bench_ukf.err; echo "ukf rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d/bench_ukf.json'))
print("UKF value %.4g ms/sweep %.1f frac %.3f kernel_ms %.3f" % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['accuracy'].get('ukf_route_instance_steps'))
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ukf_(back2|front2|eig3)_kernel' -s 4750 -c 5 -o $O/prof_ukf3 -f $U > $O/ncu_ukf3.log 2>&1
tail -2 $O/ncu_ukf3.log
ls -la $O
