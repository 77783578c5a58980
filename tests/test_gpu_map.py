"""Patch / map building on the GPU against the oracle: transform, removeOutliers, buildMap — bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import slam3d_b200
    c = slam3d_b200.Context()
    yield c
    c.close()


def pose(tx, ty, yaw):
    T = np.eye(4)
    T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    T[:3, 3] = [tx, ty, 0.01]
    return T


def test_transform_bit_exact(ctx, oracle_mod, kitti):
    T = pose(12.3, -4.5, 0.7)
    got = ctx.transform_cloud(kitti[0], T)
    want = oracle_mod.transform_cloud(kitti[0], T)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("radius,min_nb", [(0.2, 3), (0.5, 10), (0.05, 1)])
def test_remove_outliers_bit_exact(ctx, oracle_mod, kitti, radius, min_nb):
    c = kitti[1]
    got = ctx.remove_outliers(c, radius, min_nb)
    want, keep = oracle_mod.remove_outliers(c, radius, min_nb)
    assert got.shape == want.shape and 0 < got.shape[0] < c.shape[0]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # same survivors, same (input) order


def test_remove_outliers_edge_cases(ctx, kitti):
    import slam3d_b200
    c = slam3d_b200.as_xyzw(kitti[0][:5000])
    assert np.array_equal(ctx.remove_outliers(c, 0.0, 3), c)      # radius <= 0: input handed back (:214)
    assert np.array_equal(ctx.remove_outliers(c, 0.2, 0), c)      # min_neighbors == 0: same
    assert ctx.remove_outliers(np.zeros((0, 3), np.float32), 0.2, 3).shape[0] == 0
    lonely = np.array([[0, 0, 0], [10, 0, 0], [20, 0, 0]], np.float32)
    assert ctx.remove_outliers(lonely, 0.2, 1).shape[0] == 0
    dup = np.zeros((5, 3), np.float32)
    assert ctx.remove_outliers(dup, 0.2, 4).shape[0] == 5 and ctx.remove_outliers(dup, 0.2, 5).shape[0] == 0


def test_build_map_bit_exact(ctx, oracle_mod, kitti):
    poses = [pose(0.69 * i, 0.004 * i, 0.0035 * i) for i in range(4)]
    got = ctx.build_map(kitti, poses, 0.2, 3, 0.1)
    want = oracle_mod.build_map(kitti, poses, 0.2, 3, 0.1)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # no outlier removal, coarser map
    got = ctx.build_map(kitti[:2], poses[:2], 0.0, 0, 0.5)
    want = oracle_mod.build_map(kitti[:2], poses[:2], 0.0, 0, 0.5)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # reference test `map_building` (PointCloudSensorTest.cpp:71-96): an empty cloud must not throw
    assert ctx.build_map([np.zeros((0, 3), np.float32)], [np.eye(4)], 0.2, 3, 0.1).shape[0] == 0
    assert ctx.build_map([], [], 0.2, 3, 0.1).shape[0] == 0


def test_build_map_2m_points(ctx, oracle_mod):
    """16 synthetic scans (2 097 152 points) into one map: accumulate + outlier removal + 0.1 m voxel grid."""
    from slam3d_b200 import synth
    rng = np.random.default_rng(20260117)
    scene = synth.Scene(20260117)
    scans, poses = [], []
    for i in range(16):
        P = synth.make_pose([12.0 * i / 15 - 6.0, 0.0, 0.0], [0.0, 0.0, 0.02 * i])
        scans.append(synth.scan(scene, P, rng)); poses.append(P)
    got = ctx.build_map(scans, poses, 0.2, 3, 0.1)
    want = oracle_mod.build_map(scans, poses, 0.2, 3, 0.1)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_combined_measurement_bit_exact(ctx, oracle_mod, kitti):
    """createCombinedMeasurement (PointCloudSensor.cpp:258-266): accumulate in list order, then re-express in the patch frame;
    both float roundings of the two pcl::transformPointCloud calls are kept, so the result equals the oracle's bit for bit —
    and two separate transform calls of the product."""
    poses = [pose(0.69 * i, 0.004 * i, 0.0035 * i) for i in range(3)]
    patch_pose = pose(0.7, 0.01, 0.004)
    got = ctx.combined_measurement(kitti[:3], poses, patch_pose)
    want = oracle_mod.combined_measurement(kitti[:3], poses, patch_pose)
    assert got.shape == want.shape == (sum(c.shape[0] for c in kitti[:3]), 4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    inv = oracle_mod.isometry_inverse(patch_pose)
    two_step = np.concatenate([ctx.transform_cloud(ctx.transform_cloud(c, P), inv) for c, P in zip(kitti[:3], poses)])
    assert np.array_equal(got.view(np.uint32), two_step.view(np.uint32))
    assert ctx.combined_measurement([], [], patch_pose).shape[0] == 0
