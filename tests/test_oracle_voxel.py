"""The oracle's VoxelGrid restatement against an independent numpy derivation and the frozen fixtures."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def numpy_voxel(xyz, leaf):
    """Independent float32 re-derivation of PCL's leaf index and centroids."""
    xyz = np.asarray(xyz, np.float32)
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = xyz.min(0), xyz.max(0)
    d = ((mx - mn) * inv).astype(np.int64) + 1
    if int(d[0]) * int(d[1]) * int(d[2]) > 2**31 - 1:
        return None, None
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = (max_b - min_b + 1).astype(np.int64)
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int32).astype(np.int64)
    key = (ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]).astype(np.uint32)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    ends = np.r_[starts[1:], len(ks)]
    out = np.ones((len(starts), 4), np.float32)
    for j, (s, e) in enumerate(zip(starts, ends)):
        acc = np.zeros(3, np.float32)
        for i in order[s:e]:
            acc = acc + xyz[i]
        out[j, :3] = acc / np.float32(e - s)
    return key, out


@pytest.mark.parametrize("leaf", [0.1, 0.2, 1.0])
def test_voxel_matches_numpy(oracle_mod, kitti, leaf):
    c = kitti[0][:40000]
    out, leaf_index, overflow = oracle_mod.voxel_downsample(c, leaf)
    key, ref = numpy_voxel(c, leaf)
    assert not overflow
    assert np.array_equal(leaf_index, key)
    assert out.shape == ref.shape
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))  # bit-exact centroids


@pytest.mark.parametrize("leaf", ["0.05", "0.1", "0.2", "0.5", "1.0"])
def test_voxel_golden(oracle_mod, kitti, golden, leaf):
    for c, g in zip(kitti, golden["voxel"][leaf]):
        out, leaf_index, overflow = oracle_mod.voxel_downsample(c, float(leaf))
        assert out.shape[0] == g["n_out"] and overflow == g["overflow"]
        assert sha(leaf_index) == g["leaf_sha"] and sha(out) == g["out_sha"]


def test_voxel_counts_match_survey(golden):
    # SURVEY Appendix B (measured independently during the survey)
    assert [r["n_out"] for r in golden["voxel"]["0.1"]] == [60152, 59567, 58853, 58098]
    assert [r["n_out"] for r in golden["voxel"]["0.2"]] == [31834, 31481, 30882, 30435]
    assert [r["n_out"] for r in golden["voxel"]["0.05"]] == [91767, 91495, 91177, 90572]


def test_voxel_edge_cases(oracle_mod):
    out, li, ov = oracle_mod.voxel_downsample(np.zeros((0, 3), np.float32), 0.1)
    assert out.shape[0] == 0 and not ov  # PointCloudSensor.cpp:193
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    out, li, ov = oracle_mod.voxel_downsample(one, 0.1)
    assert out.shape[0] == 1 and np.allclose(out[0], [1, 2, 3, 1]) and li[0] == 0
    # int32 index overflow: PCL returns the input unchanged (A.1 step 3)
    big = np.array([[0, 0, 0], [3000.0, 3000.0, 300.0], [1, 1, 1]], np.float32)
    out, li, ov = oracle_mod.voxel_downsample(big, 0.05)
    assert ov and out.shape[0] == 3 and np.array_equal(out[:, :3], big)
    # non-finite points are skipped
    nf = np.array([[0, 0, 0], [np.nan, 0, 0], [0.01, 0.01, 0.01], [np.inf, 1, 1]], np.float32)
    out, li, ov = oracle_mod.voxel_downsample(nf, 0.1)
    assert out.shape[0] == 1 and li[1] == 0xFFFFFFFF and li[3] == 0xFFFFFFFF
    # duplicates collapse to one exact centroid
    dup = np.tile(np.array([[0.3, -0.7, 1.1]], np.float32), (7, 1))
    out, li, ov = oracle_mod.voxel_downsample(dup, 0.2)
    assert out.shape[0] == 1


def test_map_building_oracle(oracle_mod, kitti):
    """Oracle restatement of transform / RadiusOutlierRemoval / buildMap against numpy + scipy re-derivations."""
    from scipy.spatial import cKDTree
    c = kitti[0][:30000]
    T = np.eye(4); T[:3, 3] = [3.0, -1.0, 0.2]
    a = 0.4
    T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    got = oracle_mod.transform_cloud(c, T)
    x, y, z = (c[:, i].astype(np.float64) for i in range(3))
    for r in range(3):
        ref = (x * T[r, 0] + (y * T[r, 1] + (z * T[r, 2] + T[r, 3]))).astype(np.float32)
        assert np.array_equal(got[:, r], ref)
    out, keep = oracle_mod.remove_outliers(c, 0.2, 3)
    tree = cKDTree(c.astype(np.float64))
    cnt = np.array([len(v) for v in tree.query_ball_point(c.astype(np.float64), 0.2)])
    assert (keep == (cnt >= 4)).mean() > 0.9995          # float32 vs float64 distances differ only on the boundary
    assert np.array_equal(out[:, :3], c[keep])
    # buildMap == the three steps
    m = oracle_mod.build_map([c, kitti[1][:30000]], [np.eye(4), T], 0.2, 3, 0.1)
    accu = np.concatenate([oracle_mod.transform_cloud(c, np.eye(4)), oracle_mod.transform_cloud(kitti[1][:30000], T)], 0)
    filt, _ = oracle_mod.remove_outliers(accu, 0.2, 3)
    ds, _, _ = oracle_mod.voxel_downsample(filt, 0.1)
    assert np.array_equal(m, ds)
    assert oracle_mod.build_map([np.zeros((0, 3), np.float32)], [np.eye(4)], 0.2, 3, 0.1).shape[0] == 0
