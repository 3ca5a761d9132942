#!/bin/bash
# Closing visit of the re-entry session: smoke(), whole GPU suite, all-configs bench line, reference arm, ncu launch list of the bench
# command, DRAM traffic of one sweep-kernel launch (profiles/ncu_traffic.json), ncu --set full of the sweep kernel at the full tile
set -u
O=gpurun_out/${1:-r02zz}
mkdir -p $O gpurun_out/traffic
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:ekf_sweep_kernel -c 400 --csv --log-file gpurun_out/traffic/sweep.csv python scripts/traffic_probe.py sweep > gpurun_out/traffic/sweep.log 2>&1; echo "traffic rc=$?"
python scripts/traffic_collect.py > $O/traffic_collect.log 2>&1; cp profiles/ncu_traffic.json $O/ncu_traffic.json
( time timeout 1500 python bench.py --steps 20 --warmup 5 > $O/bench_all.json 2> $O/bench_all.err ) 2> $O/bench_all.time; echo "bench rc=$?"
tail -3 $O/bench_all.time | head -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-sub --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launch $O/launches.csv $O/launches.txt > /dev/null 2>&1
export SLAM_TUNE="0=50"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ekf_sweep_kernel<.int.4, .bool.0>" -s 39 -c 1 -o /tmp/prof_sweep -f python bench.py --filter ekf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_sweep.log 2>&1
unset SLAM_TUNE
python scripts/ncu_summary.py full /tmp/prof_sweep.ncu-rep $O/ekf_sweep_full.txt > /dev/null 2>&1
ncu -i /tmp/prof_sweep.ncu-rep --page source --csv > /tmp/sweep_source.csv 2>/dev/null
INNER=1 python scripts/sass_profile.py /tmp/sweep_source.csv live_ekf_slam_b200/csrc/ekf_batch.o ekf_sweep_kernelILi4ELb0 40 > $O/ekf_sweep_lines.txt 2>&1
python - "$O" <<'PY'
import json, sys
O=sys.argv[1]
d=json.load(open(O+'/bench_all.json'))
e=d['e2e']
print("EKF value %.4g e2e %.4g per_tick %.4g async %.4g roof %.3f hbm_util %.3f" % (d['value'], e['value'], e['per_tick_value'], e['per_tick_async_value'], d['roofline']['frac'], d['roofline']['step_kernel']['hbm_utilisation_model']))
for k,v in d['configs'].items():
    rf=v['roofline']
    print(k, "value %.4g ms %.1f frac %.3f" % (v['value'], v['ms_per_step'], rf['frac']), rf.get('whole_step',{}).get('frac'), (v.get('e2e') or {}).get('value'), (v.get('cpu_baseline') or {}).get('value'))
r=json.load(open(O+'/bench_ref.json')); print("reference arm %.4g updates/s on %d cores" % (r['value'], r['cpu_baseline']['cores']))
PY
head -12 $O/launches.txt
