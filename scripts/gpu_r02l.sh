#!/bin/bash
# large-map path: parity tests, then the configs[3] bench record
set -u
O=gpurun_out/r02l
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_large_map.py -m gpu -q -x > $O/pytest_large.log 2>&1; echo "pytest rc=$?" >> $O/pytest_large.log
tail -4 $O/pytest_large.log
timeout 900 python bench.py --filter large --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_large.json 2> $O/bench_large.err; echo "large rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l/bench_large.json'))
rf=d['roofline']
print("large value %.5g steps/s ms/step %.4f gemm frac %.3f whole %s" % (d['value'], d['ms_per_step'], rf['frac'], rf.get('whole_step')))
PY
