"""CPU oracle of the patch / map-building helpers (PointCloudSensor.cpp:211-233, :301-318) against independent derivations
(numpy double arithmetic, scipy cKDTree neighbour counts, the voxel oracle) and the frozen hashes in tests/golden/golden.json."""
import hashlib

import numpy as np
import pytest
from scipy.spatial import cKDTree


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pose(tx, ty, yaw):
    T = np.eye(4)
    T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    T[:3, 3] = [tx, ty, 0.01]
    return T


def np_transform(c, T):
    """pcl::transformPointCloud with a double matrix: x*c0 + (y*c1 + (z*c2 + c3)) in double, then cast to float."""
    x, y, z = (c[:, i].astype(np.float64) for i in range(3))
    out = np.empty((c.shape[0], 3), np.float32)
    for r in range(3):
        out[:, r] = (x * T[r, 0] + (y * T[r, 1] + (z * T[r, 2] + T[r, 3]))).astype(np.float32)
    return out


def test_transform_matches_numpy(oracle_mod, kitti):
    T = pose(12.3, -4.5, 0.7)
    got = oracle_mod.transform_cloud(kitti[0], T)
    assert np.array_equal(got[:, :3].view(np.uint32), np_transform(kitti[0], T).view(np.uint32))
    assert np.all(got[:, 3] == 1.0)  # pcl::PointXYZ padding
    ident = oracle_mod.transform_cloud(kitti[0][:1000], np.eye(4))
    assert np.array_equal(ident[:, :3], kitti[0][:1000])


@pytest.mark.parametrize("radius,min_nb", [(0.2, 3), (0.5, 10)])
def test_remove_outliers_matches_kdtree_counts(oracle_mod, kitti, radius, min_nb):
    """RadiusOutlierRemoval keeps a point when more than min_neighbors points (itself included) lie within the radius, in
    input order.  float32 squared distances decide in PCL; a double cKDTree agrees except on points within rounding of the
    radius, which are excluded from the comparison."""
    c = kitti[1][::3]
    out, keep = oracle_mod.remove_outliers(c, radius, min_nb)
    assert out.shape[0] == keep.sum() and np.array_equal(out[:, :3], c[keep])
    tree = cKDTree(c.astype(np.float64))
    lo = np.array([len(v) for v in tree.query_ball_point(c.astype(np.float64), radius * (1 - 1e-5))])
    hi = np.array([len(v) for v in tree.query_ball_point(c.astype(np.float64), radius * (1 + 1e-5))])
    sure = lo == hi
    assert sure.mean() > 0.99
    assert np.array_equal(keep[sure], (lo > min_nb)[sure])
    assert 0 < keep.sum() < c.shape[0]


def test_remove_outliers_edge_cases(oracle_mod):
    lonely = np.array([[0, 0, 0], [10, 0, 0], [20, 0, 0]], np.float32)
    assert oracle_mod.remove_outliers(lonely, 0.2, 1)[0].shape[0] == 0
    dup = np.zeros((5, 3), np.float32)
    assert oracle_mod.remove_outliers(dup, 0.2, 4)[0].shape[0] == 5
    assert oracle_mod.remove_outliers(dup, 0.2, 5)[0].shape[0] == 0


def test_build_map_is_accumulate_filter_downsample(oracle_mod, kitti):
    """buildMap (:301-318) = getAccumulatedCloud (transform + concatenate in list order) -> removeOutliers -> downsample."""
    clouds = [c[::2] for c in kitti[:3]]
    poses = [pose(0.69 * i, 0.004 * i, 0.0035 * i) for i in range(3)]
    got = oracle_mod.build_map(clouds, poses, 0.2, 3, 0.1)
    accu = np.concatenate([oracle_mod.transform_cloud(c, T)[:, :3] for c, T in zip(clouds, poses)], 0)
    filt, _ = oracle_mod.remove_outliers(accu, 0.2, 3)
    want, _, overflow = oracle_mod.voxel_downsample(filt[:, :3], 0.1)
    assert not overflow
    assert np.array_equal(got[:, :3].view(np.uint32), want[:, :3].view(np.uint32))
    plain = oracle_mod.build_map(clouds[:2], poses[:2], 0.0, 0, 0.5)  # no outlier removal
    accu2 = np.concatenate([oracle_mod.transform_cloud(c, T)[:, :3] for c, T in zip(clouds[:2], poses[:2])], 0)
    assert np.array_equal(plain[:, :3], oracle_mod.voxel_downsample(accu2, 0.5)[0][:, :3])


def test_map_golden(oracle_mod, kitti, golden):
    g = golden["map"]
    T = pose(12.3, -4.5, 0.7)
    assert sha(oracle_mod.transform_cloud(kitti[0], T)) == g["transform cloud1"]["sha"]
    out, keep = oracle_mod.remove_outliers(kitti[1], 0.2, 3)
    assert int(keep.sum()) == g["remove_outliers cloud2 r=0.2 n=3"]["kept"] and sha(out) == g["remove_outliers cloud2 r=0.2 n=3"]["sha"]
    poses = [pose(0.69 * i, 0.004 * i, 0.0035 * i) for i in range(4)]
    m = oracle_mod.build_map(kitti, poses, 0.2, 3, 0.1)
    assert m.shape[0] == g["build_map 4 clouds"]["n_out"] and sha(m) == g["build_map 4 clouds"]["sha"]


def test_combined_measurement_is_two_rounded_transforms(oracle_mod, kitti):
    """createCombinedMeasurement = transformPointCloud(pose_i) then transformPointCloud(patch_pose^-1), each rounding to float."""
    def pose(tx, ty, yaw):
        T = np.eye(4)
        T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        T[:3, 3] = [tx, ty, 0.01]
        return T
    clouds = [kitti[0][:2000], kitti[1][:3000]]
    poses = [pose(0.5, 0.1, 0.02), pose(1.2, -0.1, 0.05)]
    patch = pose(0.6, 0.0, 0.03)
    got = oracle_mod.combined_measurement(clouds, poses, patch)
    inv = oracle_mod.isometry_inverse(patch)
    assert np.allclose(inv @ patch, np.eye(4), atol=1e-12)
    want = []
    for c, P in zip(clouds, poses):
        p = c[:, :3].astype(np.float64)
        step1 = (p[:, 0:1] * P[:3, 0] + (p[:, 1:2] * P[:3, 1] + (p[:, 2:3] * P[:3, 2] + P[:3, 3]))).astype(np.float32).astype(np.float64)
        want.append((step1[:, 0:1] * inv[:3, 0] + (step1[:, 1:2] * inv[:3, 1] + (step1[:, 2:3] * inv[:3, 2] + inv[:3, 3]))).astype(np.float32))
    want = np.concatenate(want)
    assert got.shape == (5000, 4) and np.array_equal(got[:, :3].view(np.uint32), want.view(np.uint32)) and np.all(got[:, 3] == 1.0)
