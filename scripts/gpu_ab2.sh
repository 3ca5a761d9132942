#!/bin/bash
# A/B of whole libraries on the Monte-Carlo sweep: in-tree build vs every _ab/*.so (3 x best-of-3 each, interleaved)
set -u
O=gpurun_out/${1:-ab2}
mkdir -p $O
L=live_ekf_slam_b200/libslam_filter.so
cp $L $O/orig.so
for round in 1 2; do
for v in intree _ab/*.so; do
  n=$(basename $v .so)
  if [ "$v" != intree ]; then cp $v $L; else cp $O/orig.so $L; fi
  echo "$n: $(timeout 300 python scripts/sweep_chunks.py 0 2>/dev/null | head -1 | cut -c1-175)" | tee -a $O/ab.txt
done
done
cp $O/orig.so $L; rm -f $O/orig.so
