"""CUDA NDT branch of align() (ndt.cu; PointCloudSensor.cpp:84-117) against the CPU oracle.

Gate: pose within 1e-4 m / 1e-4 rad, fitness relative error <= 1e-4, identical status / converged.  The device runs the
same ndt_math.h as tests/test_oracle_ndt.py checks on the host; only the order of the double sums differs, so the GPU
is expected to follow the oracle's iterate sequence (same iteration and pair counts) — asserted as well."""
import numpy as np
import pytest

from conftest import pose_delta
from slam3d_b200 import _abi, synth
from slam3d_b200._abi import RegistrationParameters

pytestmark = pytest.mark.gpu
TOL_T, TOL_R, TOL_FIT = 1e-4, 1e-4, 1e-4


@pytest.fixture(scope="module")
def ctx():
    import slam3d_b200
    c = slam3d_b200.Context()
    yield c
    c.close()


def ndt_params(**kw):
    kw.setdefault("point_cloud_density", 0.2)
    return RegistrationParameters.defaults(registration_algorithm=_abi.ALG_NDT, **kw)


def check(got, want, same_sequence=True):
    assert got.status == want.status, (got.status, want.status)
    assert (got.n_source, got.n_target) == (want.n_source, want.n_target)
    if want.status == _abi.S3D_TOO_FEW_POINTS:
        return
    dt, dr = pose_delta(want.pose(), got.pose())
    assert dt < TOL_T and dr < TOL_R, (dt, dr)
    assert abs(got.fitness - want.fitness) <= TOL_FIT * max(abs(want.fitness), 1e-12)
    assert got.converged == want.converged
    if same_sequence:
        assert (got.outer_iterations, got.inner_iterations, got.n_correspondences) == (want.outer_iterations, want.inner_iterations, want.n_correspondences)


@pytest.mark.parametrize("density", [0.1, 0.2])
def test_kitti_pairs_vs_oracle(ctx, oracle_mod, kitti, density):
    p = ndt_params(point_cloud_density=density)
    for i in range(3):
        check(ctx.gicp_align(kitti[i], kitti[i + 1], None, p), oracle_mod.gicp_align(kitti[i], kitti[i + 1], None, p))


def test_synthetic_pairs_vs_oracle(ctx, oracle_mod):
    src, tgt, truth = synth.scan_pair(seed=11)
    p = ndt_params(point_cloud_density=0.1)
    got, want = ctx.gicp_align(src, tgt, None, p), oracle_mod.gicp_align(src, tgt, None, p)
    check(got, want)
    dt, dr = pose_delta(got.pose(), truth)
    assert got.status == _abi.S3D_OK and dt < 0.02 and dr < 2e-3
    guess = synth.make_pose([0.05, -0.03, 0.01], [0.002, -0.003, 0.004]) @ truth
    check(ctx.gicp_align(src, tgt, guess, p), oracle_mod.gicp_align(src, tgt, guess, p))
    # other voxel sizes / step lengths / outlier ratios, loop-closure sized motion
    src, tgt, truth = synth.scan_pair(seed=5, loop=True)
    for kw in (dict(point_cloud_density=0.5, resolution=2.0, step_size=0.2, max_translation=5.0),
               dict(point_cloud_density=0.2, resolution=0.5, outlier_ratio=0.55),
               dict(point_cloud_density=0.3, resolution=1.5, step_size=0.1, transformation_epsilon=1e-9, maximum_iterations=12)):
        p = ndt_params(**kw)
        g = truth if kw.get("resolution") == 0.5 else None
        check(ctx.gicp_align(src, tgt, g, p), oracle_mod.gicp_align(src, tgt, g, p))


def test_gates_match_oracle(ctx, oracle_mod, kitti):
    src, tgt = kitti[0][::2], kitti[1][::2]
    for kw in (dict(max_fitness_score=1e-6), dict(max_translation=0.05), dict(maximum_iterations=1), dict(maximum_iterations=3, step_size=0.2),
               dict(point_cloud_density=0.0), dict(resolution=0.01), dict(point_cloud_density=2.0, resolution=0.5)):
        p = ndt_params(**kw)
        got, want = ctx.gicp_align(src, tgt, None, p), oracle_mod.gicp_align(src, tgt, None, p)
        check(got, want)
    r = ctx.gicp_align(src, tgt, None, ndt_params(resolution=0.01))  # int32 voxel index overflow: "Voxel grid is not searchable"
    assert r.status == _abi.S3D_NOT_CONVERGED and r.outer_iterations == 0 and np.array_equal(r.pose(), np.eye(4))
    r = ctx.gicp_align(src[:500], tgt[:500], None, ndt_params(point_cloud_density=20.0))
    assert r.status == _abi.S3D_TOO_FEW_POINTS
    import slam3d_b200
    with pytest.raises(slam3d_b200.S3DError, match="resolution must be positive"):
        ctx.gicp_align(src, tgt, None, ndt_params(resolution=0.0))
    # NDT_OMP (PointCloudSensor.cpp:153-157, pclomp's multi-threaded NDT) selects the NDT branch of the GPU path
    r_omp = ctx.gicp_align(src, tgt, None, RegistrationParameters.defaults(point_cloud_density=0.2, registration_algorithm=_abi.ALG_NDT_OMP))
    r_ndt = ctx.gicp_align(src, tgt, None, ndt_params())
    assert r_omp.status == r_ndt.status and np.array_equal(r_omp.pose(), r_ndt.pose()) and r_omp.outer_iterations == r_ndt.outer_iterations


def test_batch_equals_single_and_is_deterministic(ctx, kitti):
    p = ndt_params(point_cloud_density=0.3)
    srcs = [kitti[0], kitti[1], kitti[2], kitti[0][::3], kitti[2][:90]]
    tgts = [kitti[1], kitti[2], kitti[3], kitti[1][::3], kitti[3][:90]]
    batch = ctx.gicp_align_batch(srcs, tgts, None, p)
    again = ctx.gicp_align_batch(srcs, tgts, None, p)
    for i in range(len(srcs)):
        one = ctx.gicp_align(srcs[i], tgts[i], None, p)
        for other in (batch[i], again[i]):
            assert other.status == one.status
            assert np.array_equal(other.pose(), one.pose()) and other.fitness == one.fitness
            assert (other.outer_iterations, other.inner_iterations, other.n_correspondences) == (one.outer_iterations, one.inner_iterations, one.n_correspondences)
    assert batch[4].status == _abi.S3D_TOO_FEW_POINTS


def test_host_mirror_runs_ndt(kitti, oracle_mod):
    """PointCloudSensor::createConstraint with registration_algorithm = NDT goes through the same C-ABI call."""
    import ctypes as C
    import test_gpu_host as th
    host = th.load_host()
    sensor = host.s3dhost_sensor_create(b"velodyne")
    p = ndt_params(point_cloud_density=0.3)
    host.s3dhost_sensor_set_params(sensor, C.byref(p), 0)
    I = np.eye(4)
    st, T, info, msg = th.create_constraint(host, sensor, kitti[0], kitti[1], I, I, I)
    want = oracle_mod.gicp_align(kitti[0], kitti[1], None, p)
    assert st == 0, msg
    dt, dr = pose_delta(want.pose(), T)
    assert dt < TOL_T and dr < TOL_R
    p2 = ndt_params(point_cloud_density=0.3, max_fitness_score=1e-6)
    host.s3dhost_sensor_set_params(sensor, C.byref(p2), 0)
    st, T, info, msg = th.create_constraint(host, sensor, kitti[0], kitti[1], I, I, I)
    assert st == 1 and "NDT failed with Fitness-Score" in msg
    host.s3dhost_sensor_destroy(sensor)


@pytest.mark.parametrize("density", [0.1, 0.2])
def test_kitti_pairs_vs_golden(ctx, kitti, golden, density):
    """Committed fixtures (tests/golden/golden.json "align_ndt", frozen oracle outputs on the reference's test/cloud1-4.bin)."""
    p = ndt_params(point_cloud_density=density)
    for a, b in ((0, 1), (1, 2), (2, 3)):
        g = golden["align_ndt"][f"cloud{a+1}->cloud{b+1}@{density}"]
        r = ctx.gicp_align(kitti[a], kitti[b], None, p)
        assert r.status == g["status"] and r.converged == g["converged"]
        assert (r.n_source, r.n_target) == (g["n_source"], g["n_target"])
        dt, dr = pose_delta(g["T"], r.pose())
        assert dt < TOL_T and dr < TOL_R, (dt, dr)
        assert abs(r.fitness - g["fitness"]) <= TOL_FIT * g["fitness"]
        assert (r.outer_iterations, r.inner_iterations, r.n_correspondences) == (g["outer_iterations"], g["inner_iterations"], g["n_correspondences"])
