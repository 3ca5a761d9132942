#!/bin/bash
# A/B on the GPU box: for each _ab/*.so (scripts/build_variant.sh) swap it in for the in-tree library and run "$@" (default:
# the UKF bench line), printing the value and the per-kernel times of one launch list.  The in-tree library is restored.
set -u
O=gpurun_out/ab
mkdir -p $O
L=live_ekf_slam_b200/libslam_filter.so
cp $L $O/orig.so
CMD=${@:-python bench.py --filter ukf --steps 1 --warmup 3 --no-e2e --no-cpu-baseline}
for v in base _ab/*.so; do
  n=$(basename $v .so)
  if [ "$v" != base ]; then cp $v $L; else cp $O/orig.so $L; fi
  timeout 600 $CMD > $O/$n.json 2> $O/$n.err; echo "$n rc=$?"
  python - $O/$n.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   value %.4g  ms/step %.1f  kernel_ms %.3f" % (d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_launch', 0)), d.get('accuracy', {}).get('mean_pos_err_m'), d.get('accuracy', {}).get('ukf_route_instance_steps'))
except Exception as e: print("   no line:", e)
PY
done
cp $O/orig.so $L
