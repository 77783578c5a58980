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
    hdr = os.path.join(ROOT, "slam3d_b200", "csrc", "gicp_math.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    lib.hm_f.restype = C.c_double
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


def moments(hm, a, b, M):
    p = np.concatenate([a[:, :3].astype(np.float64), np.ones((a.shape[0], 1))], 1)
    q = b[:, :3].astype(np.float64)
    mom = np.zeros(74)
    for aa in range(3):
        for bb in range(aa, 3):
            for c in range(4):
                for e in range(c, 4):
                    mom[hm.hm_sym3(aa, bb) * 10 + hm.hm_sym4(c, e)] = np.sum(M[:, aa, bb] * p[:, c] * p[:, e])
    Mq = np.einsum("nij,nj->ni", M, q)
    for aa in range(3):
        for c in range(4):
            mom[60 + aa * 4 + c] = np.sum(Mq[:, aa] * p[:, c])
    mom[72] = np.einsum("ni,ni->", q, Mq)
    mom[73] = a.shape[0]
    return mom


def oracle_objective(oracle_mod, a, b, M, x):
    f = C.c_double(0)
    g = np.zeros(6)
    H = np.zeros((6, 6), order="F")
    Mc = np.ascontiguousarray(M.transpose(0, 2, 1))
    xx = np.ascontiguousarray(x, np.float64)
    oracle_mod.lib().s3d_oracle_test_objective(ptr(a), ptr(b), ptr(Mc), a.shape[0], ptr(xx), C.byref(f), ptr(g), ptr(H))
    return f.value, g, np.array(H)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_moment_objective_matches_oracle(hm, oracle_mod, seed):
    a, b, M = make_problem(seed)
    mom = moments(hm, a, b, M)
    x = np.array([0.25, -0.08, 0.04, 0.01, -0.015, 0.02])
    f, g, H = oracle_objective(oracle_mod, a, b, M, x)
    fm = hm.hm_f(ptr(mom), ptr(x))
    gm = np.zeros(6); Hm = np.zeros((6, 6), order="F")
    hm.hm_dfddf(ptr(mom), ptr(x), ptr(gm), ptr(Hm))
    # the oracle transforms points in float32 (PCL); the moment form is exact in double: agreement ~1e-6 relative
    assert abs(fm - f) < 2e-5 * abs(f)
    assert np.allclose(gm, g, rtol=1e-4, atol=1e-4 * np.abs(g).max())
    assert np.allclose(np.array(Hm), H, rtol=1e-5, atol=1e-5 * np.abs(H).max())


@pytest.mark.parametrize("seed", [3, 4])
def test_newton_from_moments_matches_oracle(hm, oracle_mod, seed):
    a, b, M = make_problem(seed, m=2000, noise=0.02)
    mom = moments(hm, a, b, M)
    T0 = np.eye(4, dtype=np.float32, order="F")
    To = T0.copy(order="F"); Tm = T0.copy(order="F")
    io = C.c_int(0); im = C.c_int(0)
    Mc = np.ascontiguousarray(M.transpose(0, 2, 1))
    oracle_mod.lib().s3d_oracle_test_newton(ptr(a), ptr(b), ptr(Mc), a.shape[0], ptr(To), 20, C.byref(io))
    st = hm.hm_newton(ptr(mom), ptr(Tm), 20, C.byref(im))
    assert st == 0
    assert np.abs(np.array(To) - np.array(Tm)).max() < 2e-6
    assert abs(io.value - im.value) <= 1
    # fewer than 4 correspondences: PCL throws, we report failure
    mom4 = moments(hm, a[:3], b[:3], M[:3])
    assert hm.hm_newton(ptr(mom4), ptr(Tm), 20, C.byref(im)) == 2


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
