#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_sim_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ekf.json 2> gpurun_out/bench_ekf.err; tail -3 gpurun_out/bench_ekf.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ekf.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['pipelined_value'])
r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','kernel_ms_per_launch')}); s=r.get('step_kernel',r); print({k:s[k] for k in ('kernel','achieved','frac','kernel_ms_per_launch')})
PY
