#!/bin/bash
# A/B builds: scripts/build_variant.sh NAME "-DSOME_SWITCH ..." [file.cu ...]  ->  _ab/NAME.so (the listed translation units
# recompiled with the extra flags, default ukf_batch.cu; the other objects come from the regular build).  scripts/gpu_ab.sh
# swaps each _ab/*.so in for the in-tree library on the GPU box and runs the same bench command.
set -e
cd "$(dirname "$0")/.."
NAME=$1; FLAGS=$2; shift 2
FILES=${@:-ukf_batch.cu}
C=live_ekf_slam_b200/csrc
make -C $C -j8 >/dev/null
mkdir -p _ab/obj_$NAME
OBJS=""
for f in capi.cu ekf_batch.cu ekf_large.cu ukf_batch.cu sim.cu; do
  if [[ " $FILES " == *" $f "* ]]; then
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS -c -o _ab/obj_$NAME/${f%.cu}.o $C/$f
    OBJS="$OBJS _ab/obj_$NAME/${f%.cu}.o"
  else OBJS="$OBJS $C/${f%.cu}.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _ab/$NAME.so $OBJS
echo "built _ab/$NAME.so"
