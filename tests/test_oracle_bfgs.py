"""The oracle's BFGS mode (PCL <= 1.13's inner optimiser: pcl/registration/bfgs.h, a port of GSL vector_bfgs2; oracle/bfgs_oracle.inc).
Unpinned like the rest of the oracle: checked here against an independent minimiser (scipy) of the same objective, against the
Newton mode (both minimise the same function, so the converged poses must agree far inside slam3d's gates), and against the
invariants of Fletcher's line search."""
import ctypes as C

import numpy as np
import pytest

from conftest import pose_delta
from slam3d_b200._abi import RegistrationParameters
from test_oracle_gicp import _objective, _rot


def _problem(seed, m=400, noise=0.02, offset=(0.15, -0.1, 0.05, 0.02, -0.015, 0.03)):
    rng = np.random.default_rng(seed)
    a = np.ones((m, 4), np.float32); a[:, :3] = rng.uniform(-5, 5, (m, 3))
    x = np.array(offset)
    b = np.ones((m, 4), np.float32)
    b[:, :3] = (a[:, :3].astype(np.float64) @ _rot(x).T + x[:3] + rng.normal(0, noise, (m, 3))).astype(np.float32)
    L = rng.normal(size=(m, 3, 3))
    M = L @ L.transpose(0, 2, 1) + 0.5 * np.eye(3)
    return a, b, M, x


def _state_of(T):
    return np.array([T[0, 3], T[1, 3], T[2, 3], np.arctan2(T[2, 1], T[2, 2]), np.arcsin(np.clip(-T[2, 0], -1, 1)), np.arctan2(T[1, 0], T[0, 0])])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_bfgs_minimises_the_gicp_objective(oracle_mod, seed):
    """From the identity, estimateRigidTransformationBFGS must end where an independent BFGS (scipy, same f and gradient through
    the oracle's functor hook) ends: same minimum of f, gradient below PCL's tolerance, the known offset recovered."""
    from scipy.optimize import minimize
    a, b, M, truth = _problem(seed)
    Mc = np.ascontiguousarray(M.transpose(0, 2, 1))
    T, inner, st = oracle_mod.test_bfgs(a, b, Mc, np.eye(4), max_inner=50)
    assert st == 0 and 1 <= inner <= 50
    x = _state_of(T.astype(np.float64))
    f, g, _ = _objective(oracle_mod, a, b, M, x)
    ref = minimize(lambda z: _objective(oracle_mod, a, b, M, z)[0], np.zeros(6), jac=lambda z: _objective(oracle_mod, a, b, M, z)[1], method="BFGS",
                   options={"gtol": 1e-6})
    assert np.linalg.norm(g) < 1e-2                      # testGradient(gradient_tol = 1e-2)
    assert f <= ref.fun + 1e-4 * max(1.0, ref.fun)       # PCL stops at |g| < 1e-2: within that slack of the true minimum
    assert np.allclose(x, ref.x, atol=2e-3)
    assert np.allclose(x, truth, atol=5e-3)


def test_bfgs_iteration_cap_and_descent(oracle_mod):
    """maximum_optimizer_iterations caps the BFGS steps (PCL accepts the state at the cap), every accepted step satisfies
    Fletcher's rho test (sufficient decrease), and too few correspondences are refused (NotEnoughPointsException)."""
    a, b, M, _ = _problem(7, offset=(0.4, 0.3, -0.2, 0.05, 0.04, -0.06))
    Mc = np.ascontiguousarray(M.transpose(0, 2, 1))
    f_prev = _objective(oracle_mod, a, b, M, np.zeros(6), grad=False)[0]
    for cap in (1, 2, 3, 5):
        T, inner, st = oracle_mod.test_bfgs(a, b, Mc, np.eye(4), max_inner=cap)
        assert st == 0 and inner <= cap
        f = _objective(oracle_mod, a, b, M, _state_of(T.astype(np.float64)), grad=False)[0]
        assert f < f_prev * (1 + 1e-9)
        f_prev = f
    T, inner, st = oracle_mod.test_bfgs(a[:3], b[:3], Mc[:3], np.eye(4))
    assert st != 0


def test_align_bfgs_mode_agrees_with_newton_mode(oracle_mod, kitti):
    """align() with the BFGS inner optimiser (PCL <= 1.13) and with the Newton one (PCL >= 1.14, what the CUDA path mirrors) on the
    reference's own test clouds: same decisions, and poses as close as PCL's inner tolerance allows.  Both stop an inner solve at
    |g| < 1e-2; along the driving direction the objective's curvature is ~1-3 (planes constrain it weakly), so that tolerance is
    worth millimetres — and BFGS, whose first step of every solve is a line search along -g, reaches it after a shorter move than
    a Newton step: here it declares the outer loop converged at iteration 4 (3.6 mm short in x) where Newton runs to 9.  This is
    why a golden file from a BFGS-default PCL must be compared with the oracle in BFGS mode (tests/test_pcl_golden.py), not with
    the Newton results at 1e-4."""
    p = RegistrationParameters.defaults(point_cloud_density=0.2)
    newton = oracle_mod.gicp_align(kitti[0], kitti[1], None, p)
    old = oracle_mod.set_gicp_optimizer("bfgs")
    try:
        assert old == "newton"
        bf = oracle_mod.gicp_align(kitti[0], kitti[1], None, p)
    finally:
        oracle_mod.set_gicp_optimizer("newton")
    assert (bf.status, bf.converged, bf.n_source, bf.n_target) == (newton.status, newton.converged, newton.n_source, newton.n_target)
    assert (bf.status, bf.converged) == (0, 1)
    dt, dr = pose_delta(bf.pose(), newton.pose())
    assert dt < 1e-2 and dr < 1e-3, (dt, dr)
    assert abs(bf.fitness - newton.fitness) < 1e-3
    assert bf.inner_iterations >= bf.outer_iterations
    again = oracle_mod.gicp_align(kitti[0], kitti[1], None, p)  # the switch is restored
    assert np.array_equal(np.asarray(again.pose()), np.asarray(newton.pose()))
