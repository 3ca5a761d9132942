#!/bin/bash
set -u
O=gpurun_out/r01h
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_ekf_n4.json 2> $O/bench_ekf_n4.err; echo "ekf n4 rc=$?"
wc -l $O/bench_ekf_n4.json; cut -c1-260 $O/bench_ekf_n4.json
