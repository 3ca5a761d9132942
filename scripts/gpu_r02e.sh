#!/bin/bash
# Round-2 visit E: zero-copy per-tick path, triangular lm_gemm: parity, then the all-configs bench line
set -u
O=gpurun_out/r02e
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_large_map.py tests/test_gpu_sim_parity.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 > $O/bench_all.json 2> $O/bench_all.err ) 2> $O/bench_all.time; echo "bench rc=$?"
tail -3 $O/bench_all.time; tail -5 $O/bench_all.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e/bench_all.json'))
e=d['e2e']
print("EKF value %.4g e2e %.4g per_tick %.4g (%.1f us) async %.4g diff %s %s" % (d['value'], e['value'], e['per_tick_value'], e['per_tick_us'], e['per_tick_async_value'], e['per_tick_vs_replay_max_pose_diff'], e['per_tick_async_max_pose_diff']))
for k,v in d['configs'].items():
    rf=v['roofline']
    print(k, "value %.4g ms %.1f frac %.3f" % (v['value'], v['ms_per_step'], rf['frac']), {kk: rf[kk] for kk in ('kernel_ms_per_launch','gemm_share_of_step','executed_tflops') if kk in rf}, rf.get('whole_step'))
    if v.get('e2e'): print("   e2e %.4g" % v['e2e']['value'], v['e2e'].get('per_tick_value'))
PY
