"""CUDA VoxelGrid (s3d_voxel_downsample) against the CPU oracle: bit-exact leaf assignment and centroids."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    import slam3d_b200
    c = slam3d_b200.Context()
    yield c
    c.close()


@pytest.mark.parametrize("leaf", [0.05, 0.1, 0.2, 0.5, 1.0])
def test_kitti_bit_exact(ctx, oracle_mod, kitti, golden, leaf):
    for c, g in zip(kitti, golden["voxel"][str(leaf)]):
        out, li, ov = ctx.voxel_downsample(c, leaf)
        eo, eli, eov = oracle_mod.voxel_downsample(c, leaf)
        assert ov == eov and out.shape == eo.shape
        assert np.array_equal(li, eli)                                    # leaf assignment, bit exact
        assert np.array_equal(out.view(np.uint32), eo.view(np.uint32))   # centroids, bit exact, same order
        assert sha(li) == g["leaf_sha"] and sha(out) == g["out_sha"] and out.shape[0] == g["n_out"]


def test_edge_cases(ctx, oracle_mod):
    out, li, ov = ctx.voxel_downsample(np.zeros((0, 3), np.float32), 0.1)
    assert out.shape[0] == 0 and not ov
    cases = {
        "single": np.array([[1.0, 2.0, 3.0]], np.float32),
        "overflow": np.array([[0, 0, 0], [3000.0, 3000.0, 300.0], [1, 1, 1]], np.float32),
        "nonfinite": np.array([[0, 0, 0], [np.nan, 0, 0], [0.01, 0.01, 0.01], [np.inf, 1, 1], [5, 5, 5]], np.float32),
        "duplicates": np.tile(np.array([[0.3, -0.7, 1.1]], np.float32), (700, 1)),
        "allnan": np.full((5, 3), np.nan, np.float32),
        "negative": (np.random.default_rng(0).uniform(-50, -10, (5000, 3))).astype(np.float32),
        "line": np.stack([np.linspace(0, 100, 3000), np.zeros(3000), np.zeros(3000)], 1).astype(np.float32),
    }
    for name, c in cases.items():
        leaf = 0.05 if name == "overflow" else 0.2
        out, li, ov = ctx.voxel_downsample(c, leaf)
        eo, eli, eov = oracle_mod.voxel_downsample(c, leaf)
        assert ov == eov, name
        assert np.array_equal(li, eli), name
        assert out.shape == eo.shape and np.array_equal(out.view(np.uint32), eo.view(np.uint32)), name


def test_map_cloud_2m(ctx, oracle_mod):
    """BASELINE config 3: 2 097 152-point synthetic cloud, leaf 0.05 / 0.1 / 0.2."""
    from slam3d_b200 import synth
    cloud = synth.map_cloud(n_scans=16)
    assert cloud.shape[0] == 2097152
    for leaf in (0.05, 0.1, 0.2):
        out, li, ov = ctx.voxel_downsample(cloud, leaf)
        eo, eli, eov = oracle_mod.voxel_downsample(cloud, leaf)
        assert ov == eov and np.array_equal(li, eli)
        assert out.shape == eo.shape and np.array_equal(out.view(np.uint32), eo.view(np.uint32))
        # idempotence property: every output point lies in its own voxel, one point per voxel
        out2, li2, _ = ctx.voxel_downsample(out, leaf)
        assert out2.shape[0] <= out.shape[0]


def test_device_pointer_input(ctx, oracle_mod, kitti):
    import torch
    import slam3d_b200
    a = torch.from_numpy(slam3d_b200.as_xyzw(kitti[0])).cuda()
    out, li, ov = ctx.voxel_downsample(a, 0.1)
    eo, eli, _ = oracle_mod.voxel_downsample(kitti[0], 0.1)
    assert np.array_equal(li, eli) and np.array_equal(out.view(np.uint32), eo.view(np.uint32))
