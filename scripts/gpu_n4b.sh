#!/bin/bash
# 2-GPU all-configs bench line on the final code (torchrun, one rank per GPU) + the reference arm launched the same way
set -u
O=gpurun_out/r02zzz
mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_all_n4.json 2> $O/bench_all_n4.err; echo "n2 rc=$?"
wc -l $O/bench_all_n4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02zzz/bench_all_n4.json'))
print("n_gpus", d['n_gpus'], "EKF value %.4g e2e %.4g" % (d['value'], d['e2e']['value']))
for k,v in d['configs'].items(): print(k, "value %.4g %s" % (v['value'], v.get('scaling')))
PY
