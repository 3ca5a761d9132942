"""CPU check of the generation-3 tridiagonal eigensolver (live_ekf_slam_b200/csrc/eig3.cuh): the header's scalar building
blocks are compiled with g++ into a harness that emulates the kernel's thread mapping, and compared with LAPACK on
tridiagonal matrices of the UKF (Householder-reduced scale * sym(P) along oracle runs) and on adversarial ones (exact
degeneracy from freshly inserted landmarks, split matrices, Wilkinson pairs, tight clusters)."""
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("eig3") / "harness")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", os.path.join(ROOT, "tests", "eig3_host_harness.cpp"), "-o", exe])
    return exe


def run(exe, d, e, maxc=16):
    n = len(d)
    txt = f"{n} {maxc}\n" + " ".join(repr(float(v)) for v in d) + "\n" + " ".join(repr(float(v)) for v in e[: n - 1]) + "\n"
    out = subprocess.run([exe], input=txt, capture_output=True, text=True, check=True).stdout.split("\n")
    if not out[0].startswith("ok"):
        return None, None, out[0]
    lam = np.array([float(v) for v in out[1: 1 + n]])
    V = np.array([[float(v) for v in out[1 + n + i].split()] for i in range(n)])
    return lam, V, "ok"


def quality(d, e, lam, V):
    n = len(d)
    T = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
    w, Z = np.linalg.eigh(T)
    nrm = max(np.abs(T).sum(axis=0).max(), 1e-300)
    orth = np.abs(V.T @ V - np.eye(n)).max()
    res = np.abs(T @ V - V * lam).max() / nrm
    eigerr = np.abs(np.sort(lam) - w).max() / nrm
    S = (V * np.sqrt(np.maximum(lam, 1e-8))) @ V.T          # what the UKF uses: sqrt of the clipped spectrum (ukf.cpp:120,208)
    Sr = (Z * np.sqrt(np.maximum(w, 1e-8))) @ Z.T
    return orth, res, eigerr, np.abs(S - Sr).max() / max(np.abs(Sr).max(), 1e-300)


def tridiag_of(Y):
    Hh = sl.hessenberg(Y)
    return np.diag(Hh).copy(), np.append(np.diag(Hh, -1), 0.0)


def test_eig3_on_ukf_matrices(harness, oracle):
    p, lm, fwd, ang = H.config2(seed=0, steps=420, filt="ukf_slam")
    op = H.oracle_params(oracle, p)
    stream, _ = H.oracle_meas_stream(oracle, op, lm, fwd, ang, seed=5, instance=0)
    f = oracle.OracleFilter(oracle.UKF_SLAM, op, 50)
    f.init(0, 0, 0)
    checked = 0
    for t in range(len(fwd)):
        if t in (0, 1, 2, 5, 9, 20, 40, 41, 80, 150, 151, 250, 400, 419):
            P = f.cov()
            n = P.shape[0]
            scale = float(np.float32(np.float32(n) / np.float32(1 - np.float32(0.2))))
            d, e = tridiag_of(0.5 * (P + P.T) * scale)
            lam, V, st = run(harness, d, e)
            assert st == "ok", (t, st)
            orth, res, eigerr, serr = quality(d, e, lam, V)
            assert orth <= 5e-13 and res <= 5e-15 and eigerr <= 5e-15 and serr <= 5e-13, (t, n, orth, res, eigerr, serr)
            checked += 1
        f.update(fwd[t], ang[t], stream[t], oracle.STRUCTURED)
    assert checked == 14 and f.M >= 15


def test_eig3_adversarial(harness):
    rng = np.random.default_rng(0)
    cases = {}
    cases["identity"] = (np.full(40, 130.0), np.zeros(40))
    cases["part_identity"] = (np.concatenate([rng.uniform(0.1, 5, 20), np.full(12, 130.0)]),
                              np.concatenate([rng.uniform(0.1, 1, 19), [0.0], np.zeros(12)]))
    m = 10
    dw = np.abs(np.arange(-m, m + 1)).astype(float)
    cases["wilkinson21"] = (dw, np.append(np.ones(2 * m), 0.0))
    cases["glued_wilkinson"] = (np.concatenate([dw, dw]), np.concatenate([np.ones(2 * m), [1e-7], np.ones(2 * m), [0.0]]))
    cases["tight_cluster_12"] = (1 + 1e-6 * np.arange(12), np.append(1e-6 * rng.uniform(0.5, 1, 11), 0.0))
    A = rng.normal(size=(104, 104))
    cases["wishart_104"] = tridiag_of(A @ A.T / 104)
    B = rng.normal(size=(30, 30))
    Yb = np.zeros((46, 46)); Yb[:30, :30] = B @ B.T; Yb[30:, 30:] = np.eye(16) * 132.5       # eight freshly inserted landmarks
    cases["after_insertion"] = tridiag_of(Yb)
    cases["indefinite"] = tridiag_of(B @ B.T - 3.0 * np.eye(30))
    # close pairs over the whole range between "numerically orthogonal already" and "coincident": the refinement step after
    # the first Gram-Schmidt pass is taken only when a projection removed a sizeable part of a vector (eig3.cuh REFINE_BELOW)
    for gap in (1e-4, 1e-6, 1e-8, 1e-10, 1e-12, 1e-14):
        Qr, _ = np.linalg.qr(rng.normal(size=(24, 24)))
        base = np.sort(rng.uniform(0.5, 40.0, 12))
        ev = np.concatenate([base, base + gap * 40.0])
        cases["pairs_gap_%g" % gap] = tridiag_of((Qr * ev) @ Qr.T)
    cases["zero_matrix"] = (np.zeros(7), np.zeros(7))
    cases["n_is_1"] = (np.array([2.5]), np.zeros(1))
    for name, (d, e) in cases.items():
        lam, V, st = run(harness, np.asarray(d, float), np.asarray(e, float))
        assert st == "ok", (name, st)
        orth, res, eigerr, serr = quality(np.asarray(d, float), np.asarray(e, float), lam, V)
        assert orth <= 1e-12 and res <= 1e-14 and eigerr <= 1e-14 and serr <= 1e-12, (name, orth, res, eigerr, serr)


def test_eig3_declines_large_clusters(harness):
    """clusters beyond maxc are not re-orthogonalised in the kernel: the instance is handed to the QL route"""
    rng = np.random.default_rng(1)
    d, e = 1 + 1e-6 * np.arange(30), np.append(1e-6 * rng.uniform(0.5, 1, 29), 0.0)
    assert run(harness, d, e, maxc=16)[2].startswith("declined")
    lam, V, st = run(harness, d, e, maxc=32)
    assert st == "ok" and quality(d, e, lam, V)[0] <= 1e-12
    assert run(harness, *tridiag_of(np.diag([1.0, 1.0 + 1e-9, 5.0]) + 1e-3), maxc=1)[2].startswith("declined")


def test_eig3_eigenvector_formulations_agree_bitwise(harness):
    """twisted_vector (two work vectors, second backward sweep), twisted_vector_pf (batched read-backs, parked backward pivots),
    twisted_vector1 (one work vector: the shared-memory tile) and twisted_vector1g (tile + parked pivots) are re-schedulings of
    the same arithmetic: same twist index, same norm, same components and kept factors, bit for bit."""
    rng = np.random.default_rng(5)
    mats = []
    A = rng.normal(size=(104, 104)); mats.append(tridiag_of(A @ A.T / 104))
    B = rng.normal(size=(37, 37)); mats.append(tridiag_of(B @ B.T - 2.0 * np.eye(37)))
    m = 10
    dw = np.abs(np.arange(-m, m + 1)).astype(float)
    mats.append((np.concatenate([dw, dw]), np.concatenate([np.ones(2 * m), [1e-7], np.ones(2 * m), [0.0]])))
    mats.append((1 + 1e-6 * np.arange(12), np.append(1e-6 * rng.uniform(0.5, 1, 11), 0.0)))
    for d, e in mats:
        n = len(d)
        txt = f"{n} 16\n" + " ".join(repr(float(v)) for v in d) + "\n" + " ".join(repr(float(v)) for v in e[: n - 1]) + "\n"
        out = subprocess.run([harness, "variants"], input=txt, capture_output=True, text=True, check=True).stdout
        assert out.strip().endswith("variants ok"), out
