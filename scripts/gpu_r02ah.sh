#!/bin/bash
# storage-order rank-2 walk (in-tree) vs the paired-row walk (_ab/base.so): EKF parity of the in-tree build, whole-sweep rates, per-chunk times
set -u
O=gpurun_out/r02ah; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_param_variants.py -m gpu -q -x > $O/pytest_ekf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ekf.log
tail -3 $O/pytest_ekf.log
L=live_ekf_slam_b200/libslam_filter.so
cp $L $O/orig.so
for round in 1 2; do
for v in intree _ab/*.so; do
  n=$(basename $v .so)
  if [ "$v" != intree ]; then cp $v $L; else cp $O/orig.so $L; fi
  SLAM_DEBUG_SWEEP=1 timeout 300 python scripts/sweep_chunks.py 0 > $O/chunks_$n.txt 2> $O/chunks_$n.err
  echo "$n: $(head -1 $O/chunks_$n.txt)"
done
done
cp $O/orig.so $L; rm -f $O/orig.so
