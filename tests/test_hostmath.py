"""slam3d_b200/csrc/gicp_math.h (the FP64 optimiser math that runs on the GPU) compiled for the host and checked
against the oracle: moment-form objective/gradient/Hessian, Newton step, covariance normal, Mahalanobis matrix."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(ROOT, "tests", "hostmath.cpp")
    out = os.path.join(ROOT, "tests", "_hostmath.so")
    hdrs = [os.path.join(ROOT, "slam3d_b200", "csrc", h) for h in ("gicp_math.h", "ndt_math.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def make_problem(seed, m=300, noise=0.05):
    rng = np.random.default_rng(seed)
    a = np.ones((m, 4), np.float32); a[:, :3] = rng.uniform(-30, 30, (m, 3))
    b = np.ones((m, 4), np.float32); b[:, :3] = a[:, :3] + rng.normal(0, noise, (m, 3)) + [0.3, -0.1, 0.05]
    n1 = rng.normal(size=(m, 3)); n1 /= np.linalg.norm(n1, axis=1, keepdims=True)
    n2 = rng.normal(size=(m, 3)); n2 /= np.linalg.norm(n2, axis=1, keepdims=True)
    k = 1 - 1e-3
    Cs = 2 * np.eye(3) - k * n1[:, :, None] * n1[:, None, :] - k * n2[:, :, None] * n2[:, None, :]
    M = np.linalg.inv(Cs)
    return a, b, M


def m6(M):
    return np.ascontiguousarray(np.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], 1))


def oracle_objective(oracle_mod, a, b, M, x):
    f = C.c_double(0)
    g = np.zeros(6)
    H = np.zeros((6, 6), order="F")
    Mc = np.ascontiguousarray(M.transpose(0, 2, 1))
    xx = np.ascontiguousarray(x, np.float64)
    oracle_mod.lib().s3d_oracle_test_objective(ptr(a), ptr(b), ptr(Mc), a.shape[0], ptr(xx), C.byref(f), ptr(g), ptr(H))
    return f.value, g, np.array(H)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_objective_from_sums_matches_oracle(hm, oracle_mod, seed):
    """Same float residuals, sums in a different order: f, g, H agree to double rounding."""
    a, b, M = make_problem(seed)
    x = np.array([0.25, -0.08, 0.04, 0.01, -0.015, 0.02])
    f, g, H = oracle_objective(oracle_mod, a, b, M, x)
    fm = C.c_double(0); gm = np.zeros(6); Hm = np.zeros((6, 6), order="F")
    hm.hm_objective(ptr(a), ptr(b), ptr(m6(M)), a.shape[0], ptr(x), C.byref(fm), ptr(gm), ptr(Hm))
    assert abs(fm.value - f) < 1e-12 * abs(f)
    assert np.allclose(gm, g, rtol=1e-10, atol=1e-12 * np.abs(g).max())
    assert np.allclose(np.array(Hm), H, rtol=1e-10, atol=1e-12 * np.abs(H).max())


@pytest.mark.parametrize("seed,noise", [(3, 0.02), (4, 0.02), (5, 0.3), (6, 1e-4)])
def test_newton_state_machine_reproduces_oracle(hm, oracle_mod, seed, noise):
    """The resumable optimiser follows the oracle's estimateRigidTransformationNewton decision for decision:
    bit-identical float transform and the same number of inner iterations."""
    a, b, M = make_problem(seed, m=2000, noise=noise)
    T0 = np.eye(4, dtype=np.float32, order="F")
    To = T0.copy(order="F"); Tm = T0.copy(order="F")
    io = C.c_int(0); im = C.c_int(0); ev = C.c_int(0)
    Mc = np.ascontiguousarray(M.transpose(0, 2, 1))
    oracle_mod.lib().s3d_oracle_test_newton(ptr(a), ptr(b), ptr(Mc), a.shape[0], ptr(To), 20, C.byref(io))
    st = hm.hm_newton(ptr(a), ptr(b), ptr(m6(M)), a.shape[0], ptr(Tm), 20, C.byref(im), C.byref(ev))
    assert st == 0
    assert np.array_equal(np.array(To), np.array(Tm))
    assert io.value == im.value and ev.value >= im.value + 1
    # second call starting from the result (near the optimum: the float-rounding dominated regime)
    To2 = To.copy(order="F"); Tm2 = Tm.copy(order="F")
    oracle_mod.lib().s3d_oracle_test_newton(ptr(a), ptr(b), ptr(Mc), a.shape[0], ptr(To2), 20, C.byref(io))
    hm.hm_newton(ptr(a), ptr(b), ptr(m6(M)), a.shape[0], ptr(Tm2), 20, C.byref(im), C.byref(ev))
    assert np.array_equal(np.array(To2), np.array(Tm2)) and io.value == im.value
    # fewer than 4 correspondences: PCL throws, we report failure
    assert hm.hm_newton(ptr(a), ptr(b), ptr(m6(M)), 3, ptr(Tm), 20, C.byref(im), C.byref(ev)) == 2


def test_newton_direction_indefinite(hm):
    rng = np.random.default_rng(5)
    for trial in range(20):
        A = rng.normal(size=(6, 6)); A = A + A.T
        if trial % 2 == 0:
            A = A @ A.T + 0.1 * np.eye(6)  # positive definite -> Cholesky path
        g = rng.normal(size=6)
        w, V = np.linalg.eigh(A)
        inv = np.where(w < 0, 1.0 / w.max(), 1.0 / w)
        expect = V @ (inv * (V.T @ g))
        d = np.zeros(6)
        hm.hm_direction(ptr(np.asfortranarray(A)), ptr(g), ptr(d))
        assert np.allclose(d, expect, rtol=1e-9, atol=1e-10)


def test_normal_and_mahalanobis(hm, oracle_mod, kitti):
    f = oracle_mod.voxel_downsample(kitti[0][:30000], 0.2)[0]
    idx, d2, cov = oracle_mod.knn_covariances(f, 20)
    rng = np.random.default_rng(6)
    for i in rng.choice(f.shape[0], 200, replace=False):
        nb = f[idx[i], :3]
        mean = nb.astype(np.float64).mean(0)
        c = (nb[:, :, None] * nb[:, None, :]).astype(np.float64).mean(0) - np.outer(mean, mean)
        n = np.zeros(3)
        hm.hm_normal(ptr(np.asfortranarray(c)), ptr(n))
        Creg = np.eye(3) - (1 - 1e-3) * np.outer(n, n)
        w = np.linalg.eigvalsh(c)
        if abs(w[1]) < 4 * abs(w[0]):
            continue
        assert np.allclose(Creg, cov[i], atol=1e-6)
    R = np.eye(3) + 1e-3 * rng.normal(size=(3, 3))
    n1 = np.array([0.0, 0.6, 0.8]); n2 = np.array([1.0, 0.0, 0.0])
    k = 1 - 1e-3
    C1 = np.eye(3) - k * np.outer(n1, n1); C2 = np.eye(3) - k * np.outer(n2, n2)
    Mref = np.linalg.inv(R @ C1 @ R.T + C2)
    M6 = np.zeros(6)
    hm.hm_mahalanobis(ptr(np.asfortranarray(R @ R.T)), ptr(R @ n1), ptr(n2), ptr(M6))
    assert np.allclose(M6, [Mref[0, 0], Mref[0, 1], Mref[0, 2], Mref[1, 1], Mref[1, 2], Mref[2, 2]], rtol=1e-10)
