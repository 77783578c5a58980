"""The oracle's exact NN / kNN / covariance restatement against independent implementations."""
import hashlib

import numpy as np
import pytest
from scipy.spatial import cKDTree


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def filtered(oracle_mod, kitti):
    return [oracle_mod.voxel_downsample(c, 0.1)[0] for c in kitti[:2]]


def test_kdtree_equals_bruteforce(oracle_mod, filtered):
    ref = filtered[0]
    q = ref[::37]
    bi, bd = oracle_mod.knn_bruteforce(ref, q, 20)
    ki, kd, _ = oracle_mod.knn_covariances(ref, 20)
    assert np.array_equal(ki[::37], bi)
    assert np.array_equal(kd[::37].view(np.uint32), bd.view(np.uint32))
    assert np.all(ki[:, 0] == np.arange(ref.shape[0]))  # self first (d = 0)


def test_nn_against_ckdtree(oracle_mod, filtered):
    ref, qry = filtered
    idx, d2 = oracle_mod.nearest_neighbors(ref, qry)
    dd, ii = cKDTree(ref[:, :3].astype(np.float64)).query(qry[:, :3].astype(np.float64), k=1)
    differ = np.flatnonzero(idx != ii)
    # float32 vs float64 ranking may differ only on near-ties
    assert differ.size <= 3
    for i in differ:
        assert abs(np.sqrt(d2[i]) - dd[i]) < 1e-5
    bi, bd = oracle_mod.knn_bruteforce(ref, qry[:2000], 1)
    assert np.array_equal(bi[:, 0], idx[:2000]) and np.array_equal(bd[:, 0].view(np.uint32), d2[:2000].view(np.uint32))


def test_ties_go_to_lowest_index(oracle_mod):
    # lattice with duplicated points: many exact float ties
    g = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ref = np.concatenate([g, g[::-1]], 0)  # every point twice
    idx, d2 = oracle_mod.nearest_neighbors(ref, g + np.float32(0.5))
    bi, bd = oracle_mod.knn_bruteforce(ref, g + np.float32(0.5), 1)
    assert np.array_equal(idx, bi[:, 0])
    n = g.shape[0]
    d = ((ref[None, :, :3] - (g[:, None, :] + np.float32(0.5))) ** 2).sum(-1)
    assert np.array_equal(idx, np.argmin(d, 1))  # argmin returns the first (lowest) index of a tie
    ki, kd, _ = oracle_mod.knn_covariances(ref, 8)
    bi, bd = oracle_mod.knn_bruteforce(ref, ref, 8)
    assert np.array_equal(ki, bi) and np.array_equal(kd, bd)
    assert n == 216


def test_transform_order_of_queries(oracle_mod, filtered):
    ref, qry = filtered
    T = np.eye(4)
    T[:3, 3] = [0.68, 0.003, 0.007]
    a = 0.0031
    T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    Tf = T.astype(np.float32)
    x, y, z = qry[:, 0], qry[:, 1], qry[:, 2]
    q = np.ones_like(qry)
    for r in range(3):  # ((c0*x + c1*y) + c2*z) + c3 in float32
        q[:, r] = ((Tf[r, 0] * x + Tf[r, 1] * y) + Tf[r, 2] * z) + Tf[r, 3]
    i1, d1 = oracle_mod.nearest_neighbors(ref, qry, T)
    i2, d2 = oracle_mod.nearest_neighbors(ref, q)
    assert np.array_equal(i1, i2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))


def test_covariances_against_numpy(oracle_mod, filtered):
    ref = filtered[0]
    idx, d2, cov = oracle_mod.knn_covariances(ref, 20)
    rng = np.random.default_rng(0)
    for i in rng.choice(ref.shape[0], 300, replace=False):
        nb = ref[idx[i], :3]
        mean = nb.astype(np.float64).mean(0)
        prod = (nb[:, :, None] * nb[:, None, :]).astype(np.float64)  # float32 products, double accumulate
        c = prod.mean(0) - np.outer(mean, mean)
        w, v = np.linalg.eigh(c)
        order = np.argsort(-np.abs(w))
        if abs(w[order[1]]) < 4 * abs(w[order[2]]):
            continue  # ill-conditioned normal (SURVEY hard part 3)
        n = v[:, order[2]]
        expect = np.eye(3) - (1 - 1e-3) * np.outer(n, n)
        assert np.allclose(cov[i], expect, atol=1e-6)
        ev = np.linalg.eigvalsh(cov[i])
        assert np.allclose(ev, [1e-3, 1, 1], atol=1e-9)


def test_eigen_solvers(oracle_mod):
    import ctypes as C
    rng = np.random.default_rng(1)
    lib = oracle_mod.lib()
    for n, fn in ((3, lib.s3d_oracle_test_eigen3), (6, lib.s3d_oracle_test_eigen6)):
        for _ in range(50):
            a = rng.normal(size=(n, n))
            a = a + a.T
            A = np.asfortranarray(a)
            V = np.zeros((n, n), order="F")
            w = np.zeros(n)
            fn(A.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
            assert np.allclose(V @ np.diag(w) @ V.T, a, atol=1e-12)
            assert np.allclose(V.T @ V, np.eye(n), atol=1e-12)
            assert np.allclose(np.sort(w), np.linalg.eigvalsh(a), atol=1e-12)


def test_knn_golden(oracle_mod, filtered, golden):
    idx, d2, cov = oracle_mod.knn_covariances(filtered[0], 20)
    g = golden["knn"]["cloud1@0.1,k=20"]
    assert sha(idx) == g["index_sha"] and sha(d2) == g["dist2_sha"]
    assert abs(cov.sum() - g["cov_sum"]) < 1e-6
    i, d = oracle_mod.nearest_neighbors(filtered[0], filtered[1])
    g = golden["knn"]["nn cloud2@0.1 -> cloud1@0.1"]
    assert sha(i) == g["index_sha"] and sha(d) == g["dist2_sha"]
