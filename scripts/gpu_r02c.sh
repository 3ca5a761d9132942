#!/bin/bash
# Round-2 visit C: generation 3 of the UKF step (parallel eigensolver + dense products): parity, then the UKF bench line
set -u
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ukf_parity.py tests/test_gpu_param_variants.py tests/test_gpu_loc_naive.py tests/test_golden.py -m gpu -q -x > $O/pytest_ukf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ukf.log
tail -15 $O/pytest_ukf.log
timeout 900 python bench.py --filter ukf --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_ukf.json 2> $O/bench_ukf.err; echo "ukf rc=$?"
tail -3 $O/bench_ukf.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c/bench_ukf.json'))
print("UKF value %.4g ms/sweep %.1f frac %.3f kernel_ms %.3f" % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['accuracy']['mean_pos_err_m'], d['accuracy']['bad_instances'], d['accuracy'].get('ukf_route_instance_steps'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 33 --csv --log-file $O/ukf_launches_mid.csv python bench.py --filter ukf --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --filter-steps 1000 --warmup 0 > /dev/null 2> $O/ncu_mid.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 10500 -c 33 --csv --log-file $O/ukf_launches.csv python bench.py --filter ukf --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --filter-steps 1000 --warmup 0 > /dev/null 2> $O/ncu.err
python - <<'PY'
import csv, collections
for f in ('ukf_launches_mid.csv', 'ukf_launches.csv'):
    print('==', f)
    rows=[r for r in csv.reader(open('gpurun_out/r02c/'+f)) if len(r)>5]
    hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
    h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
    agg=collections.defaultdict(list)
    for r in rows[hdr+1:]:
        try: agg[r[ki][:60]].append(float(r[vi].replace(',','')))
        except: pass
    for k,v in agg.items():
        if sum(v)/len(v) > 20e3: print("%-62s n=%3d mean %.1f us" % (k,len(v),sum(v)/len(v)/1e3))
PY
