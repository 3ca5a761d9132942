#!/bin/bash
set -u
O=gpurun_out/r02ai; mkdir -p $O
SLAM_DEBUG_SWEEP=1 timeout 600 python scripts/sweep_chunks.py 0 32 64 96 128 > $O/chunks.txt 2> $O/chunks.err
cut -c1-60 $O/chunks.txt
