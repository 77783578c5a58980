"""CUDA exact kNN / covariances / 1-NN (s3d_knn_covariances, s3d_nearest_neighbors) against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import slam3d_b200
    c = slam3d_b200.Context()
    yield c
    c.close()


@pytest.fixture(scope="module")
def filtered(oracle_mod, kitti):
    return [oracle_mod.voxel_downsample(c, 0.1)[0] for c in kitti[:2]]


def test_knn_indices_bit_exact(ctx, oracle_mod, filtered):
    for f in filtered:
        gi, gd, gc = ctx.knn_covariances(f, 20)
        oi, od, oc = oracle_mod.knn_covariances(f, 20)
        assert np.array_equal(gi, oi)
        assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
        # the regularised covariance is rebuilt from the same normal: identical up to rounding of U diag U^T
        bad = np.abs(gc - oc).max(axis=(1, 2)) > 1e-9
        assert bad.mean() < 2e-3, bad.sum()  # exactly degenerate neighbourhoods may pick another in-plane direction


@pytest.mark.parametrize("k", [1, 5, 20, 32, 50, 128, 200, 201, 300])  # > 200: the heap moves from shared to global memory
def test_knn_other_k(ctx, oracle_mod, filtered, k):
    f = filtered[0][::5] if k <= 128 else filtered[0][::25]
    gi, gd, _ = ctx.knn_covariances(f, k)
    oi, od, _ = oracle_mod.knn_covariances(f, k)
    assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_knn_ties_and_duplicates(ctx, oracle_mod):
    g = np.stack(np.meshgrid(np.arange(7), np.arange(7), np.arange(7), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ref = np.concatenate([g, g[::-1], g[::3]], 0)  # lattice with duplicated points: exact float ties everywhere
    gi, gd, _ = ctx.knn_covariances(ref, 20)
    oi, od = oracle_mod.knn_bruteforce(ref, ref, 20)
    assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_knn_sparse_far_field(ctx, oracle_mod):
    rng = np.random.default_rng(5)
    dense = rng.uniform(-2, 2, (3000, 3))
    far = rng.uniform(-400, 400, (200, 3))  # isolated points: the search has to climb to the top level
    ref = np.concatenate([dense, far], 0).astype(np.float32)
    gi, gd, _ = ctx.knn_covariances(ref, 20)
    oi, od = oracle_mod.knn_bruteforce(ref, ref, 20)
    assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_nn_bit_exact(ctx, oracle_mod, filtered):
    ref, qry = filtered
    T = np.eye(4); T[:3, 3] = [0.68, 0.003, 0.007]
    a = 0.0031
    T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    for tf in (None, T):
        gi, gd = ctx.nearest_neighbors(ref, qry, tf)
        oi, od = oracle_mod.nearest_neighbors(ref, qry, tf)
        assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_nn_queries_outside_and_ties(ctx, oracle_mod):
    rng = np.random.default_rng(6)
    ref = rng.uniform(-10, 10, (5000, 3)).astype(np.float32)
    qry = np.concatenate([rng.uniform(-300, 300, (2000, 3)), rng.uniform(-10, 10, (2000, 3))], 0).astype(np.float32)
    gi, gd = ctx.nearest_neighbors(ref, qry)
    oi, od = oracle_mod.knn_bruteforce(ref, qry, 1)
    assert np.array_equal(gi, oi[:, 0]) and np.array_equal(gd.view(np.uint32), od[:, 0].view(np.uint32))
    g = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ref = np.concatenate([g, g[::-1]], 0)
    gi, gd = ctx.nearest_neighbors(ref, g + np.float32(0.5))
    oi, od = oracle_mod.knn_bruteforce(ref, g + np.float32(0.5), 1)
    assert np.array_equal(gi, oi[:, 0])


def test_tiny_clouds(ctx, oracle_mod):
    ref = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5]], np.float32)
    gi, gd, _ = ctx.knn_covariances(ref, 3)
    oi, od = oracle_mod.knn_bruteforce(ref, ref, 3)
    assert np.array_equal(gi, oi)
    same = np.zeros((50, 3), np.float32)
    gi, gd, _ = ctx.knn_covariances(same, 4)
    oi, od = oracle_mod.knn_bruteforce(same, same, 4)
    assert np.array_equal(gi, oi)


def test_map_cloud_2m_covariances(ctx, oracle_mod):
    """BASELINE config 3: kNN-20 + covariance on the voxel-filtered 2 097 152-point synthetic map cloud."""
    from slam3d_b200 import synth
    cloud = synth.map_cloud(n_scans=16)
    for leaf in (0.1, 0.2):
        f, _, _ = ctx.voxel_downsample(cloud, leaf, want_leaf_index=False)
        gi, gd, gc = ctx.knn_covariances(f, 20)
        rng = np.random.default_rng(int(leaf * 100))
        pick = rng.choice(f.shape[0], 1500, replace=False)
        oi, od = oracle_mod.knn_bruteforce(f, f[pick], 20)
        assert np.array_equal(gi[pick], oi) and np.array_equal(gd[pick].view(np.uint32), od.view(np.uint32))
        # size-independent properties on the full result: self first, distances ascending, unit-trace structure of C
        assert np.array_equal(gi[:, 0], np.arange(f.shape[0], dtype=np.uint32)) or np.all(gd[:, 0] == 0)
        assert np.all(np.diff(gd, axis=1) >= 0)
        ev = np.linalg.eigvalsh(gc[::997])
        assert np.allclose(ev, [1e-3, 1, 1], atol=1e-9)


def test_sparse_cloud_overflows_hash_arena_and_retries(ctx, oracle_mod):
    """Every point alone in its cell on many levels: the optimistic hash arena overflows and the library re-runs the
    batch with the size the device reported (no error, exact results)."""
    rng = np.random.default_rng(9)
    g = np.stack(np.meshgrid(np.arange(22), np.arange(22), np.arange(22), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ref = (g + rng.uniform(-0.2, 0.2, g.shape)).astype(np.float32)
    gi, gd, _ = ctx.knn_covariances(ref, 8)
    pick = rng.choice(ref.shape[0], 800, replace=False)
    oi, od = oracle_mod.knn_bruteforce(ref, ref[pick], 8)
    assert np.array_equal(gi[pick], oi) and np.array_equal(gd[pick].view(np.uint32), od.view(np.uint32))
    ni, nd = ctx.nearest_neighbors(ref, ref[pick] + np.float32(0.3))
    bi, bd = oracle_mod.knn_bruteforce(ref, ref[pick] + np.float32(0.3), 1)
    assert np.array_equal(ni, bi[:, 0]) and np.array_equal(nd.view(np.uint32), bd[:, 0].view(np.uint32))
