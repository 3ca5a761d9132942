#!/bin/bash
# closing ncu --set full captures (source counters) of the round's five hot kernels on the final code; the reports are
# summarised on the box (metrics + per-source-line stall samples) and only the text comes back
set -u
O=gpurun_out/r02m
mkdir -p $O
cap() {  # name, kernel regex, skip, object, command...
  local name=$1 kern=$2 skip=$3 obj=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -o /tmp/prof_$name -f "$@" > $O/ncu_$name.log 2>&1
  python scripts/ncu_summary.py full /tmp/prof_$name.ncu-rep $O/${name}_full.txt > /dev/null 2>&1
  python scripts/sass_lines.py /tmp/prof_$name.ncu-rep $obj $kern 30 > $O/${name}_lines.txt 2>&1
  head -3 $O/${name}_lines.txt
  rm -f /tmp/prof_$name.ncu-rep
}
U="python bench.py --filter ukf --steps 1 --warmup 0 --filter-steps 1000 --no-e2e --no-cpu-baseline"
L="python bench.py --filter large --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
C=live_ekf_slam_b200/csrc
# launch indices: filter step 900 with two batch slices (4 eig3 tile/hand-over launches, 3 front2 classes, 2 back3 widths per slice)
cap eig3 ukf_eig3_kernel 7202 $C/ukf_batch.o $U
cap front2 ukf_front2_kernel 5402 $C/ukf_batch.o $U
cap back3 ukf_back3_kernel 3600 $C/ukf_batch.o $U
cap lm_front lm_front 2800 $C/ekf_large.o $L
cap lm_gemm lm_gemm 2800 $C/ekf_large.o $L
