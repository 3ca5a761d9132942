#!/bin/bash
# 2-GPU check of the launch contract: torchrun, one rank per GPU; stdout must be exactly one JSON line
set -u
O=gpurun_out/r01g
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_ekf_n2.json 2> $O/bench_ekf_n2.err; echo "ekf n2 rc=$?"
wc -l $O/bench_ekf_n2.json; head -c 80 $O/bench_ekf_n2.json; echo; grep -c "NCCL version" $O/bench_ekf_n2.err
env | grep -i nccl
