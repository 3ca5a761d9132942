#!/bin/bash
# Round-2 visit G: DRAM traffic of one launch per dominant kernel (with the algorithmic bytes of the same launch) + sweep chunk probe
set -u
mkdir -p gpurun_out/traffic
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:ekf_step_kernel -c 4000 --csv --log-file gpurun_out/traffic/step.csv python scripts/traffic_probe.py step > gpurun_out/traffic/step.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:ekf_sweep_kernel -c 400 --csv --log-file gpurun_out/traffic/sweep.csv python scripts/traffic_probe.py sweep > gpurun_out/traffic/sweep.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:'ukf_(front2|eig3|back3)_kernel' -c 4000 --csv --log-file gpurun_out/traffic/ukf.csv python scripts/traffic_probe.py ukf > gpurun_out/traffic/ukf.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:lm_gemm -c 4000 --csv --log-file gpurun_out/traffic/gemm.csv python scripts/traffic_probe.py gemm > gpurun_out/traffic/gemm.log 2>&1
for f in gpurun_out/traffic/*.log; do tail -n 1 $f | cut -c1-200; done
