#!/bin/bash
set -u
O=gpurun_out/${1:-r02ab}
mkdir -p $O
L=live_ekf_slam_b200/libslam_filter.so
cp $L $O/orig.so
for v in intree _ab/*.so; do
  n=$(basename $v .so)
  if [ "$v" != intree ]; then cp $v $L; else cp $O/orig.so $L; fi
  SLAM_DEBUG_SWEEP=1 timeout 300 python scripts/sweep_chunks.py 0 128 > $O/chunks_$n.txt 2> $O/chunks_$n.err
  echo "== $n"; cut -c1-60 $O/chunks_$n.txt
done
cp $O/orig.so $L; rm -f $O/orig.so
