#!/bin/bash
# visit r02s: whole GPU suite on the new sweep launcher (chunk 48, width by tile size, known-ID-only core in the sweep kernel) + EKF bench line
set -u
O=gpurun_out/r02s
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python scripts/tune_sweep.py 2>&1 | head -6 | tee $O/tune.txt
timeout 600 python bench.py --filter ekf --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_ekf.json 2> $O/bench_ekf.err
python -c "
import json; d=json.load(open('$O/bench_ekf.json')); e=d['e2e']; print('bench value %.5g ms %.2f e2e %.5g per_tick %.4g async %.4g' % (d['value'], d['ms_per_step'], e['value'], e['per_tick_value'], e['per_tick_async_value']), d['accuracy']['rmse_x'], d['roofline']['launches_timed'], d['gpu_launches'])"
