#!/bin/bash
# Closing visit: smoke(), the whole GPU suite, the all-configs bench line, the reference arm, the ncu launch list of the bench command
set -u
O=gpurun_out/${1:-r02z}
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 > $O/bench_all.json 2> $O/bench_all.err ) 2> $O/bench_all.time; echo "bench rc=$?"
tail -3 $O/bench_all.time | head -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-sub --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python - "$O" <<'PY'
import json, sys, csv, collections
O=sys.argv[1]
d=json.load(open(O+'/bench_all.json'))
e=d['e2e']
print("EKF value %.4g e2e %.4g per_tick %.4g async %.4g roof %.3f hbm_util %.3f" % (d['value'], e['value'], e['per_tick_value'], e['per_tick_async_value'], d['roofline']['frac'], d['roofline']['step_kernel']['hbm_utilisation_model']))
for k,v in d['configs'].items():
    rf=v['roofline']
    print(k, "value %.4g ms %.1f frac %.3f" % (v['value'], v['ms_per_step'], rf['frac']), rf.get('whole_step',{}).get('frac'), (v.get('e2e') or {}).get('value'), (v.get('cpu_baseline') or {}).get('value'))
r=json.load(open(O+'/bench_ref.json')); print("reference arm %.4g updates/s on %d cores" % (r['value'], r['cpu_baseline']['cores']))
rows=[r for r in csv.reader(open(O+'/launches.csv')) if len(r)>5]
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[hi+1:]:
    try: agg[r[ki][:48]].append(float(r[vi].replace(',','')))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:8]: print("%-50s n=%4d share %.1f%% mean %.1f us" % (k,len(v),100*sum(v)/tot,sum(v)/len(v)/1e3))
PY
