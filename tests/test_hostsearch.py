"""The product's search logic (slam3d_b200/csrc/nn_search.cuh, knn_walk.cuh: what the per-thread walk scans, prunes and
certifies, the 4-ary heap, the two-pass 27-block) compiled for the host — one host thread plays one device thread
(tests/cuda_host_shim.h, tests/hostsearch.cpp) — and checked against brute force without a GPU.  Same parity contract as the
GPU tests: flann::L2_Simple<float> distances, exact search, ties to the lowest index, bit-exact."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
NO_INDEX = 0xFFFFFFFF


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(ROOT, "tests", "hostsearch.cpp")
    out = os.path.join(ROOT, "tests", "_hostsearch.so")
    deps = [src, os.path.join(ROOT, "tests", "cuda_host_shim.h")] + [os.path.join(ROOT, "slam3d_b200", "csrc", h) for h in ("common.cuh", "nn_search.cuh", "knn_walk.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "tests"),
                               "-o", out, src])
    lib = C.CDLL(out)
    lib.hs_build_grid.restype = C.c_void_p
    lib.hs_build_grid.argtypes = [C.c_void_p, C.c_uint64, C.c_float]
    lib.hs_free_grid.argtypes = [C.c_void_p]
    lib.hs_levels.argtypes = [C.c_void_p]
    lib.hs_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_int] + [C.c_void_p] * 5
    lib.hs_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Grid:
    def __init__(self, lib, cloud, leaf_hint):
        self.lib = lib
        self.cloud = np.ascontiguousarray(cloud[:, :3], np.float32)
        self.h = lib.hs_build_grid(ptr(self.cloud), self.cloud.shape[0], leaf_hint)

    def nn(self, queries, cutoff2=np.inf, gather=False, hints=None):
        q = np.ascontiguousarray(queries[:, :3], np.float32)
        n = q.shape[0]
        idx = np.empty(n, np.uint32); d2 = np.empty(n, np.float32); lb2 = np.empty(n, np.float32); pos = np.empty(n, np.uint32)
        hp = ptr(np.ascontiguousarray(hints, np.uint32)) if hints is not None else None
        self.lib.hs_nn(self.h, ptr(q), n, np.float32(cutoff2), int(gather), hp, ptr(idx), ptr(d2), ptr(lb2), ptr(pos))
        return idx, d2, lb2, pos

    def knn(self, k):
        n = self.cloud.shape[0]
        idx = np.empty((n, k), np.uint32); d2 = np.empty((n, k), np.float32)
        self.lib.hs_knn(self.h, k, ptr(idx), ptr(d2))
        return idx, d2

    def close(self):
        self.lib.hs_free_grid(self.h)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def filtered(oracle_mod, kitti):
    return [oracle_mod.voxel_downsample(c, 0.2)[0][:, :3] for c in kitti[:2]]


@pytest.mark.parametrize("gather", [False, True])
def test_nn_lidar_bit_exact(hs, oracle_mod, filtered, gather):
    ref, qry = filtered
    g = Grid(hs, ref, 0.2)
    shift = qry + np.float32([0.68, 0.003, 0.007])  # a typical odometry offset: distances of 0 .. 1 m
    for q in (qry, shift):
        idx, d2, lb2, _ = g.nn(q, gather=gather)
        oi, od = oracle_mod.knn_bruteforce(ref, q, 2)
        assert np.array_equal(idx, oi[:, 0]) and np.array_equal(bits(d2), bits(od[:, 0]))
        # lb2 bounds every OTHER point from below (the temporal-coherence certificate of gicp_iter_kernel relies on it)
        assert np.all(lb2 <= od[:, 1] * np.float32(1.000001))
    g.close()


def test_nn_two_pass_against_single_pass(hs, oracle_mod, filtered):
    """scan_block<true> against scan_block<false>: identical winner (index, distance, sorted position); lb2 may differ (cells
    beyond the list's capacity are scanned earlier) but stays a valid bound for every other point in both."""
    ref, qry = filtered
    g = Grid(hs, ref, 0.2)
    q = qry + np.float32([0.4, -0.2, 0.05])
    a = g.nn(q, gather=False)
    b = g.nn(q, gather=True)
    for j in (0, 1, 3):
        assert np.array_equal(bits(a[j]), bits(b[j]))
    second = oracle_mod.knn_bruteforce(ref, q, 2)[1][:, 1]
    assert np.all(a[2] <= second * np.float32(1.000001)) and np.all(b[2] <= second * np.float32(1.000001))
    assert (bits(a[2]) == bits(b[2])).mean() > 0.7  # mostly the very same scans
    # with the previous result as hint (outer iterations 2..n) and with a finite cut-off (max_correspondence_distance^2)
    moved = q + np.float32([0.05, 0.02, -0.01])
    a = g.nn(moved, cutoff2=6.25, gather=False, hints=a[3])
    b = g.nn(moved, cutoff2=6.25, gather=True, hints=b[3])
    for j in (0, 1, 3):
        assert np.array_equal(bits(a[j]), bits(b[j]))
    g.close()


def test_nn_hints_and_cutoff_do_not_change_the_result(hs, oracle_mod, filtered):
    ref, qry = filtered
    g = Grid(hs, ref, 0.2)
    rng = np.random.default_rng(3)
    q = qry[::3] + np.float32([0.3, 0.1, 0.0])
    oi, od = oracle_mod.knn_bruteforce(ref, q, 1)
    hints = rng.integers(0, ref.shape[0], q.shape[0]).astype(np.uint32)  # arbitrary (bad) hints only steer the walk
    for gather in (False, True):
        idx, d2, _, _ = g.nn(q, gather=gather, hints=hints)
        assert np.array_equal(idx, oi[:, 0]) and np.array_equal(bits(d2), bits(od[:, 0]))
        idx, d2, _, _ = g.nn(q, cutoff2=6.25, gather=gather)  # the cut-off may stop the walk early, but only beyond 2.5 m
        near = od[:, 0] < 6.25
        assert near.mean() > 0.99
        assert np.array_equal(idx[near], oi[near, 0]) and np.array_equal(bits(d2[near]), bits(od[near, 0]))
    g.close()


def test_nn_outside_queries_ties_and_sparse_clouds(hs, oracle_mod):
    rng = np.random.default_rng(6)
    ref = rng.uniform(-10, 10, (5000, 3)).astype(np.float32)
    qry = np.concatenate([rng.uniform(-300, 300, (1500, 3)), rng.uniform(-10, 10, (1500, 3))], 0).astype(np.float32)
    g = Grid(hs, ref, 0.0)  # no voxel filter: finest cell = span / 1024
    for gather in (False, True):
        idx, d2, _, _ = g.nn(qry, gather=gather)
        oi, od = oracle_mod.knn_bruteforce(ref, qry, 1)
        assert np.array_equal(idx, oi[:, 0]) and np.array_equal(bits(d2), bits(od[:, 0]))
    g.close()
    lat = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ref = np.concatenate([lat, lat[::-1]], 0)  # every point twice: exact float ties, the lowest index has to win
    g = Grid(hs, ref, 0.0)
    for gather in (False, True):
        idx, _, _, _ = g.nn(lat + np.float32(0.5), gather=gather)
        assert np.array_equal(idx, oracle_mod.knn_bruteforce(ref, lat + np.float32(0.5), 1)[0][:, 0])
    g.close()


@pytest.mark.parametrize("k", [1, 5, 20, 50])
def test_knn_walk_bit_exact(hs, oracle_mod, filtered, k):
    f = filtered[0][::3] if k != 20 else filtered[0]
    g = Grid(hs, f, 0.2)
    idx, d2 = g.knn(k)
    oi, od = oracle_mod.knn_bruteforce(f, f, k)
    assert np.array_equal(idx, oi) and np.array_equal(bits(d2), bits(od))
    g.close()


def test_knn_walk_far_field_ties_and_small_clouds(hs, oracle_mod):
    rng = np.random.default_rng(5)
    ref = np.concatenate([rng.uniform(-2, 2, (3000, 3)), rng.uniform(-400, 400, (200, 3))], 0).astype(np.float32)  # has to climb to the top level
    g = Grid(hs, ref, 0.0)
    assert hs.hs_levels(g.h) == 10
    idx, d2 = g.knn(20)
    oi, od = oracle_mod.knn_bruteforce(ref, ref, 20)
    assert np.array_equal(idx, oi) and np.array_equal(bits(d2), bits(od))
    g.close()
    lat = np.stack(np.meshgrid(np.arange(7), np.arange(7), np.arange(7), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ref = np.concatenate([lat, lat[::-1], lat[::3]], 0)  # duplicated lattice: exact ties everywhere
    g = Grid(hs, ref, 0.0)
    idx, d2 = g.knn(20)
    oi, od = oracle_mod.knn_bruteforce(ref, ref, 20)
    assert np.array_equal(idx, oi) and np.array_equal(bits(d2), bits(od))
    g.close()
    tiny = rng.uniform(-1, 1, (7, 3)).astype(np.float32)  # k > cloud size: FLANN clamps, the rest of the list stays empty
    g = Grid(hs, tiny, 0.0)
    idx, d2 = g.knn(20)
    oi, _ = oracle_mod.knn_bruteforce(tiny, tiny, 7)
    assert np.array_equal(idx[:, :7], oi) and np.all(idx[:, 7:] == NO_INDEX) and np.all(np.isinf(d2[:, 7:]))
    g.close()

