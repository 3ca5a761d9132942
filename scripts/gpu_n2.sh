#!/bin/bash
# 2-GPU check of the launch contract: torchrun, one rank per GPU, NCCL all-reduce of the statistics vector / histogram.
set -u
O=gpurun_out/r01g
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.csv
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_ekf_n2.json 2> $O/bench_ekf_n2.err; echo "ekf n2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --filter mixed --instances 1024 --filter-steps 500 --steps 2 --warmup 3 > $O/bench_mixed_n2.json 2> $O/bench_mixed_n2.err; echo "mixed n2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus 2 --filter ukf --instances 4096 --filter-steps 300 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_ukf_n2.json 2> $O/bench_ukf_n2.err; echo "ukf n2 rc=$?"
cat $O/bench_mixed_n2.json | cut -c1-400; cat $O/bench_ukf_n2.json | cut -c1-300; tail -2 $O/bench_mixed_n2.err
