#!/bin/bash
# visit r02y: did the out-of-line wrap_2pi move the UKF / large-map / mixed records?
set -u
O=gpurun_out/r02y
mkdir -p $O
for f in ukf large mixed; do
  timeout 900 python bench.py --filter $f --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_$f.json 2> $O/bench_$f.err
  python -c "
import json; d=json.load(open('$O/bench_$f.json')); print('$f value %.5g %s ms %.2f' % (d['value'], d['unit'], d['ms_per_step']))"
done
