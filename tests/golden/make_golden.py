"""Generates the committed golden fixtures from the C oracle (oracle/slam_oracle.c).

The reference ships no golden vectors and cannot be run here (PARITY UNPINNED), so these are restatement outputs:
they pin the oracle against accidental drift and let the GPU tests compare against fixed numbers.
Run:  python tests/golden/make_golden.py   (rewrites tests/golden/*.npz deterministically)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_c as oc  # noqa: E402
from tests import helpers as H  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(kind, workload, steps, seed, instance, known=True, mode=None):
    p, lm, fwd, ang = workload
    p.landmark_id_is_known = known
    op = H.oracle_params(oc, p)
    stream, truth = H.oracle_meas_stream(oc, op, lm, fwd, ang, seed=seed, instance=instance)
    f = oc.OracleFilter(kind, op, 50)
    f.init(0, 0, 0)
    poses, Ms, assoc = [], [], []
    for t in range(steps):
        f.update(fwd[t], ang[t], stream[t], oc.DENSE if mode is None else mode)
        x = f.state()
        poses.append(x[:4].copy() if kind == oc.UKF_SLAM else np.r_[x[:3], 0.0])
        Ms.append(f.M)
        assoc.append(np.r_[f.assoc_log(), -9 * np.ones(8, dtype=np.int32)][:8])
    flat = np.concatenate([m.reshape(-1) for m in stream]) if steps else np.zeros(0, np.float32)
    counts = np.array([len(m) for m in stream], dtype=np.int32)
    return dict(lm=lm, fwd=fwd, ang=ang, meas_flat=flat.astype(np.float32), meas_counts=counts, truth=truth,
                poses=np.asarray(poses), Ms=np.asarray(Ms, dtype=np.int32), assoc=np.asarray(assoc, dtype=np.int32),
                x_final=f.state(), P_final=f.cov(), ids_final=f.landmark_ids(), seed=seed, instance=instance,
                known=int(known))


def main():
    np.savez_compressed(os.path.join(HERE, "ekf_config1.npz"), **run(oc.EKF_SLAM, H.config1(seed=0, steps=400), 400, 0, 0))
    np.savez_compressed(os.path.join(HERE, "ekf_unknown_ids.npz"),
                        **run(oc.EKF_SLAM, H.config2(seed=1, steps=250), 250, 3, 2, known=False))
    np.savez_compressed(os.path.join(HERE, "ukf_grid.npz"),
                        **run(oc.UKF_SLAM, H.config2(seed=2, steps=200, filt="ukf_slam"), 200, 5, 1))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
