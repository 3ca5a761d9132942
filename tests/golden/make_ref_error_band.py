"""Collects the average-position-error samples the reference recorded for its own EKF runs
(`ekf_ws/src/base_pkg/data/ekf_*/ekf.csv`, written by plotting_node.py:128 from compute_average_error,
plotting_node.py:195-218) into tests/golden/ref_ekf_avg_err.json.  These are the only numbers the reference
ships for the filter path (it has no tests); they are DATA produced by the reference's Eigen EKF + simulator and
pin the oracle statistically, not bitwise (the runs' seeds and exact settings were not recorded).

Run in the build container (needs /root/reference):  python tests/golden/make_ref_error_band.py
"""
import glob
import json
import os

REF = "/root/reference/ekf_ws/src/base_pkg/data"
out = {}
for d in sorted(glob.glob(os.path.join(REF, "ekf_*"))):
    with open(os.path.join(d, "ekf.csv")) as f:
        out[os.path.basename(d)] = [float(x) for x in f.read().split()]
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_ekf_avg_err.json")
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print(dst, {k: len(v) for k, v in out.items()})
