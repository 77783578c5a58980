"""The oracle's GICP / align() restatement (PointCloudSensor.cpp:52-82,119-174)."""
import ctypes as C

import numpy as np
import pytest

from conftest import pose_delta
from slam3d_b200 import _abi
from slam3d_b200._abi import RegistrationParameters


def _objective(oracle_mod, a, b, M, x, grad=True):
    lib = oracle_mod.lib()
    f = C.c_double(0)
    g = np.zeros(6)
    H = np.zeros((6, 6), order="F")
    Mc = np.ascontiguousarray(M.transpose(0, 2, 1))  # column-major per matrix
    x = np.ascontiguousarray(x, np.float64)
    lib.s3d_oracle_test_objective(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), Mc.ctypes.data_as(C.c_void_p),
                                  a.shape[0], x.ctypes.data_as(C.c_void_p), C.byref(f),
                                  g.ctypes.data_as(C.c_void_p) if grad else None, H.ctypes.data_as(C.c_void_p) if grad else None)
    return f.value, g, np.array(H)


def _rot(x):
    cr, sr, cp, sp, cy, sy = np.cos(x[3]), np.sin(x[3]), np.cos(x[4]), np.sin(x[4]), np.cos(x[5]), np.sin(x[5])
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def test_objective_gradient_hessian(oracle_mod):
    """f, df, ddf of gicp.hpp's OptimizationFunctorWithIndices against a float64 numpy model + finite differences."""
    rng = np.random.default_rng(3)
    m = 200
    a = np.ones((m, 4), np.float32); a[:, :3] = rng.uniform(-3, 3, (m, 3))
    b = np.ones((m, 4), np.float32); b[:, :3] = a[:, :3] + rng.normal(0, 0.05, (m, 3))
    L = rng.normal(size=(m, 3, 3))
    M = L @ L.transpose(0, 2, 1) + 0.1 * np.eye(3)
    x0 = np.array([0.1, -0.05, 0.02, 0.03, -0.02, 0.05])

    def f64(x):
        d = a[:, :3].astype(np.float64) @ _rot(x).T + x[:3] - b[:, :3].astype(np.float64)
        return np.einsum("ni,nij,nj->", d, M, d) / m

    f, g, H = _objective(oracle_mod, a, b, M, x0)
    assert abs(f - f64(x0)) < 1e-5 * max(1.0, abs(f))  # float transform inside the oracle
    eps = 1e-5
    gn = np.array([(f64(x0 + eps * e) - f64(x0 - eps * e)) / (2 * eps) for e in np.eye(6)])
    assert np.allclose(g, gn, rtol=1e-4, atol=1e-5)
    Hn = np.zeros((6, 6))
    for i, ei in enumerate(np.eye(6)):
        for j, ej in enumerate(np.eye(6)):
            h = 1e-4
            Hn[i, j] = (f64(x0 + h * ei + h * ej) - f64(x0 + h * ei - h * ej) - f64(x0 - h * ei + h * ej) + f64(x0 - h * ei - h * ej)) / (4 * h * h)
    assert np.allclose(H, Hn, rtol=1e-3, atol=1e-3)
    assert np.allclose(H, H.T)


def test_align_golden(oracle_mod, kitti, golden):
    p = RegistrationParameters.defaults(point_cloud_density=0.1)
    r = oracle_mod.gicp_align(kitti[0], kitti[1], None, p)
    g = golden["align"]["cloud1->cloud2@0.1"]
    assert (r.status, r.converged, r.outer_iterations, r.n_source, r.n_target) == (g["status"], g["converged"], g["outer_iterations"], g["n_source"], g["n_target"])
    dt, dr = pose_delta(r.pose(), g["T"])
    assert dt < 1e-9 and dr < 1e-9 and abs(r.fitness - g["fitness"]) < 1e-12


def test_align_matches_survey_proxy(golden):
    """independent float64 Gauss-Newton GICP proxy written during the survey (not PCL)."""
    proxy = {"cloud1->cloud2@0.1": ([0.6832, 0.0029, 0.0068], 0.0511),
             "cloud2->cloud3@0.1": ([0.6965, 0.0065, 0.0009], 0.0577),
             "cloud3->cloud4@0.1": ([0.7197, 0.0062, -0.0009], 0.0596)}
    for key, (t, fit) in proxy.items():
        T = np.array(golden["align"][key]["T"])
        assert np.allclose(T[:3, 3], t, atol=2e-4)
        assert abs(golden["align"][key]["fitness"] - fit) < 2e-4
        assert golden["align"][key]["status"] == 0


def test_align_recovers_synthetic_motion(oracle_mod):
    from slam3d_b200 import synth
    src, tgt, truth = synth.scan_pair(seed=7)
    p = RegistrationParameters.defaults(point_cloud_density=0.2)
    r = oracle_mod.gicp_align(src, tgt, None, p)
    assert r.status == 0 and r.converged
    dt, dr = pose_delta(truth, r.pose())
    assert dt < 0.02 and dr < 3e-3  # 2 cm range noise, 0.2 m voxels


def test_align_gates(oracle_mod, kitti):
    src, tgt = kitti[0][::4], kitti[1][::4]
    p = RegistrationParameters.defaults(point_cloud_density=0.5)
    ok = oracle_mod.gicp_align(src, tgt, None, p)
    assert ok.status == _abi.S3D_OK
    # <100 points after filtering  (PointCloudSensor.cpp:134-135)
    r = oracle_mod.gicp_align(src[:500], tgt[:500], None, RegistrationParameters.defaults(point_cloud_density=20.0))
    assert r.status == _abi.S3D_TOO_FEW_POINTS
    # fitness gate (:74-77)
    r = oracle_mod.gicp_align(src, tgt, None, RegistrationParameters.defaults(point_cloud_density=0.5, max_fitness_score=1e-6))
    assert r.status == _abi.S3D_NOT_CONVERGED and r.fitness > 1e-6
    # distance-from-guess gate (:167-172)
    r = oracle_mod.gicp_align(src, tgt, None, RegistrationParameters.defaults(point_cloud_density=0.5, max_translation=0.1))
    assert r.status == _abi.S3D_TOO_FAR_FROM_GUESS
    # algorithm switch (:139-165)
    for alg in (_abi.ALG_ICP, _abi.ALG_GICP_OMP, _abi.ALG_NDT_OMP, 17):
        r = oracle_mod.gicp_align(src, tgt, None, RegistrationParameters.defaults(point_cloud_density=0.5, registration_algorithm=alg))
        assert r.status == _abi.S3D_UNKNOWN_ALGORITHM
    # density <= 0 skips the filter (:127)
    r = oracle_mod.gicp_align(src[::8], tgt[::8], None, RegistrationParameters.defaults(point_cloud_density=0.0))
    assert r.n_source == src[::8].shape[0] and r.n_target == tgt[::8].shape[0]
    # maximum_iterations reached still counts as converged (A.4 note)
    r = oracle_mod.gicp_align(src, tgt, None, RegistrationParameters.defaults(point_cloud_density=0.5, maximum_iterations=1))
    assert r.converged == 1 and r.outer_iterations == 1


def test_align_guess_and_batch(oracle_mod, kitti):
    src, tgt = kitti[0][::4], kitti[1][::4]
    p = RegistrationParameters.defaults(point_cloud_density=0.5)
    base = oracle_mod.gicp_align(src, tgt, None, p)
    guess = np.eye(4); guess[:3, 3] = [0.6, 0.0, 0.0]
    g = oracle_mod.gicp_align(src, tgt, guess, p)
    dt, dr = pose_delta(base.pose(), g.pose())
    assert g.status == 0 and dt < 5e-3 and dr < 1e-3
    rs = oracle_mod.gicp_align_batch([src, src], [tgt, tgt], [None, guess], p)
    assert np.array_equal(rs[0].pose(), base.pose()) and np.array_equal(rs[1].pose(), g.pose())
