#!/bin/bash
# visit r02r: chunk length / headroom grid with and without the balanced grid; quick EKF bench with e2e
set -u
O=gpurun_out/r02r
mkdir -p $O
echo "== balanced grid"; timeout 600 python scripts/tune_sweep.py 2>&1 | tee $O/tune_balanced.txt
echo "== unbalanced"; SLAM_SWEEP_NO_BALANCE=1 timeout 600 python scripts/tune_sweep.py 2>&1 | head -3 | tee $O/tune_unbalanced.txt
timeout 600 python bench.py --filter ekf --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_ekf.json 2> $O/bench_ekf.err
python -c "
import json; d=json.load(open('$O/bench_ekf.json')); e=d['e2e']; print('bench value %.5g ms %.2f e2e %.5g per_tick %.4g async %.4g' % (d['value'], d['ms_per_step'], e['value'], e['per_tick_value'], e['per_tick_async_value']), d['accuracy']['rmse_x'])"
