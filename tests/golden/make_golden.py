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


def run_loc(workload, steps, seed, instance):
    """localisation-only UKF (FilterChoice::UKF_LOC): the message stream of one vehicle against the true map"""
    p, lm, fwd, ang = workload
    op = H.oracle_params(oc, p)
    stream, truth = H.oracle_meas_stream(oc, op, lm, fwd, ang, seed=seed, instance=instance)
    f = oc.OracleFilter(oc.UKF_LOC, op, 50)
    f.init(0, 0, 0)
    f.set_map(lm)
    poses, assoc = [], []
    for t in range(steps):
        f.update(fwd[t], ang[t], stream[t], oc.DENSE)
        poses.append(f.state().copy())
        assoc.append(np.r_[f.assoc_log(), -9 * np.ones(8, dtype=np.int32)][:8])
    flat = np.concatenate([m.reshape(-1) for m in stream])
    counts = np.array([len(m) for m in stream], dtype=np.int32)
    return dict(lm=lm, fwd=fwd, ang=ang, meas_flat=flat.astype(np.float32), meas_counts=counts, truth=truth,
                poses=np.asarray(poses), assoc=np.asarray(assoc, dtype=np.int32), x_final=f.state(), P_final=f.cov(),
                seed=seed, instance=instance)


def run_tsp(seed, instances, T):
    """generate_trajectory (sim_node.py:63-152) for a few Monte-Carlo instances on the 5x10 grid"""
    from live_ekf_slam_b200 import workload as wl
    p = H.Params()
    op = H.oracle_params(oc, p)
    lm = wl.grid_map_5x10()
    fw, an = [], []
    for i in instances:
        f, a = oc.tsp_trajectory(op, lm, p.landmark_noise, p.visitation_threshold, p.map_bound, p.init_pose, T, seed, i)
        fw.append(f); an.append(a)
    return dict(lm=lm, fwd=np.asarray(fw), ang=np.asarray(an), seed=seed, instances=np.asarray(instances, dtype=np.int32), T=T)


def main_new():
    """fixtures added after the first set (written separately so the older files keep their bytes)"""
    np.savez_compressed(os.path.join(HERE, "ukf_loc_grid.npz"), **run_loc(H.config2(seed=5, steps=150, filt="ukf_slam"), 150, 9, 4))
    np.savez_compressed(os.path.join(HERE, "tsp_trajectories.npz"), **run_tsp(31, [0, 3, 200], 400))


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
    if "--new" in sys.argv:
        main_new()
    else:
        main()
        main_new()
