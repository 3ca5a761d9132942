#!/bin/bash
# visit r02p: strips vs flat rank-2 walk (both with immediate barrier ids): parity of the in-tree build, per-chunk times per CTA width for both
set -u
O=gpurun_out/r02p
mkdir -p $O
L=live_ekf_slam_b200/libslam_filter.so
timeout 900 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_param_variants.py -m gpu -q -x > $O/pytest_ekf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ekf.log
tail -3 $O/pytest_ekf.log
echo "== strips (in-tree)"
SLAM_DEBUG_SWEEP=1 timeout 600 python scripts/sweep_chunks.py 0 32 64 96 128 > $O/chunks_strips.txt 2> $O/chunks_strips.err
cat $O/chunks_strips.txt
cp $L $O/orig.so
cp _ab/r2flat.so $L
echo "== flat"
SLAM_DEBUG_SWEEP=1 timeout 600 python scripts/sweep_chunks.py 0 32 64 96 128 > $O/chunks_flat.txt 2> $O/chunks_flat.err
cat $O/chunks_flat.txt
cp $O/orig.so $L
rm -f $O/orig.so
