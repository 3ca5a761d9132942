#!/bin/bash
# Round-2 visit F: the all-configs bench line under torchrun on 2 GPUs (strong split of the mixed sweep, NCCL all-reduces)
set -u
O=gpurun_out/r02g4
mkdir -p $O
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_all_n4.json 2> $O/bench_all_n4.err ) 2> $O/bench_all_n4.time; echo "bench rc=$?"
tail -3 $O/bench_all_n4.time; tail -5 $O/bench_all_n4.err; wc -l $O/bench_all_n4.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02g4/bench_all_n4.json') if l.startswith('{')][-1])
print("N", d['n_gpus'], "EKF value %.4g e2e %.4g per_tick %.4g" % (d['value'], d['e2e']['value'], d['e2e']['per_tick_value']))
for k,v in d['configs'].items():
    print(k, "value %.4g ms %.1f scaling %s frac %.3f" % (v['value'], v['ms_per_step'], v['scaling'][:8], v['roofline']['frac']), v['config'].get('instances_this_rank'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29528 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 --filter-steps 100 2>/dev/null | tail -1 | cut -c1-300
