#!/bin/bash
# per-step launch path (the kernel behind slam_step_io): forced CTA widths
set -u
O=gpurun_out/r02ad; mkdir -p $O
for t in 0 64 128 256 512; do python bench.py --filter ekf --no-sweep --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --cta-threads $t 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('cta', $t, 'value %.4g ms %.2f step-kernel ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['step_kernel']['kernel_ms_per_launch'] if 'step_kernel' in d['roofline'] else -1))"; done | tee $O/step_widths.txt
