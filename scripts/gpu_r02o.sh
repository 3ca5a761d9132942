#!/bin/bash
# visit r02o: immediate barrier ids + super-block rank-2 sweep: EKF parity, per-chunk times per CTA width, quick bench
set -u
O=gpurun_out/${1:-r02o}
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_param_variants.py tests/test_gpu_sim_parity.py -m gpu -q -x > $O/pytest_ekf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ekf.log
tail -3 $O/pytest_ekf.log
SLAM_DEBUG_SWEEP=1 timeout 600 python scripts/sweep_chunks.py ${WIDTHS:-0 32 64 96 128} > $O/chunks.txt 2> $O/chunks.err
cat $O/chunks.txt
timeout 300 python bench.py --filter ekf --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_ekf.json 2> $O/bench_ekf.err
python -c "
import json; d=json.load(open('$O/bench_ekf.json')); print('bench value %.5g ms %.2f' % (d['value'], d['ms_per_step']), d['accuracy']['rmse_x'], d['accuracy']['mean_pos_err_m'])"
