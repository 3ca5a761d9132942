#!/bin/bash
# A/B of slam_tune settings through the unmodified bench: scripts/gpu_tune_ab.sh "11=1" "11=2" ...  (each: one UKF bench line)
set -u
O=gpurun_out/tune
mkdir -p $O
for t in "$@"; do
  n=$(echo "$t" | tr '=,' '__')
  SLAM_TUNE="$t" timeout 600 python bench.py --filter ${FILTER:-ukf} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/$n.json 2> $O/$n.err
  python - $O/$n.json "$t" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-12s value %.5g  ms/step %.1f" % (sys.argv[2], d['value'], d['ms_per_step']), d.get('accuracy', {}).get('mean_pos_err_m'))
except Exception as e: print(sys.argv[2], "no line:", e)
PY
done
