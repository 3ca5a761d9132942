#!/bin/bash
# ncu full capture (source counters) of the large-map walk kernel lm_front at a steady-state step
set -u
O=gpurun_out/r02k
mkdir -p $O
U="python bench.py --filter large --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lm_front -s 2800 -c 1 -o $O/prof_lm_front -f $U > $O/ncu.log 2>&1
tail -2 $O/ncu.log
