#!/bin/bash
# 2-GPU check of the launch contract: torchrun, one rank per GPU, NCCL all-reduce of the statistics vector.
set -u
O=gpurun_out/r01e
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.csv
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_ekf_n2.json 2> $O/bench_ekf_n2.err; echo "n2 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_ekf_n1.json 2> $O/bench_ekf_n1.err; echo "n1 rc=$?"
cat $O/bench_ekf_n2.json $O/bench_ekf_n1.json; tail -3 $O/bench_ekf_n2.err
