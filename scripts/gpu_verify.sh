#!/bin/bash
# last look at the committed build: smoke() and the EKF / parameter-variant / simulator parity files
set -u
O=gpurun_out/verify; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python -m pytest tests/test_gpu_ekf_parity.py tests/test_gpu_param_variants.py tests/test_gpu_sim_parity.py tests/test_gpu_cpp_host.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 120 python scripts/sweep_chunks.py 0 2>/dev/null | head -1 | cut -c1-60
