#!/bin/bash
# visit r02q: two blocks in flight in the rank-2 walk (in-tree) vs one (variant), both with the tile-size policy for the CTA width
set -u
O=gpurun_out/r02q
mkdir -p $O
L=live_ekf_slam_b200/libslam_filter.so
timeout 900 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_param_variants.py -m gpu -q -x > $O/pytest_ekf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ekf.log
tail -3 $O/pytest_ekf.log
echo "== flat2 (in-tree)"
SLAM_DEBUG_SWEEP=1 timeout 600 python scripts/sweep_chunks.py 0 64 128 > $O/chunks_flat2.txt 2> $O/chunks_flat2.err
cat $O/chunks_flat2.txt
cp $L $O/orig.so
cp _ab/r2flat.so $L
echo "== flat"
SLAM_DEBUG_SWEEP=1 timeout 600 python scripts/sweep_chunks.py 0 64 128 > $O/chunks_flat.txt 2> $O/chunks_flat.err
cat $O/chunks_flat.txt
cp $O/orig.so $L
rm -f $O/orig.so
