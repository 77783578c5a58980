"""The PCL pin.  baseline/doicp_driver.cpp runs slam3d's downsample + doICP<pcl::GeneralizedIterativeClosestPoint>
(PointCloudSensor.cpp:52-82, 119-174, 190-201) with a real PCL on test/cloud1-4.bin and writes tests/golden/pcl_<version>.json.
When such a file is present these tests compare it with the oracle (CPU) and with the CUDA path (GPU):
  * VoxelGrid output (count + hash of the xyz floats), kNN-20 and 1-NN indices / squared distances: bit-exact;
  * align: same status / converged, pose within 1e-4 m / 1e-4 rad, fitness within 1e-4 relative (BASELINE.json north_star) —
    with the oracle in the inner-optimiser mode of that PCL ("inner_optimizer": "bfgs" up to 1.13, "newton" from 1.14 on).  The
    CUDA path mirrors the Newton optimiser only: against a BFGS golden it is held to what PCL's own inner tolerance is worth
    (1e-2 m / 1e-3 rad, tests/test_oracle_bfgs.py), against a Newton golden to 1e-4.
PCL is not installed in this repository's environment, so no such file is committed and every test here SKIPS with the reason
"parity unpinned" — which is the honest status of the oracle."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import pose_delta
from slam3d_b200._abi import RegistrationParameters

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(os.environ.get("S3D_PCL_GOLDEN_DIR", os.path.join(HERE, "golden")), "pcl_*.json")))
SKIP = "parity unpinned: no PCL golden (tests/golden/pcl_<version>.json); build baseline/doicp_driver where PCL >= 1.8.1 exists"


def cases():
    return [pytest.param(f, id=os.path.basename(f)) for f in FILES] or [pytest.param(None, id="no-pcl-golden")]


def _check_voxel_knn(pcl, impl, kitti, fnv):
    report = {}
    for leaf, rows in pcl["voxel"].items():
        for i, row in enumerate(rows):
            out = impl.voxel_downsample(kitti[i], float(leaf))[0]
            same = out.shape[0] == row["n_out"] and fnv(np.ascontiguousarray(out[:, :3])) == row["xyz_fnv1a"]
            report[f"voxel {leaf} cloud{i + 1}"] = same
    f1 = impl.voxel_downsample(kitti[0], 0.1)[0]
    f2 = impl.voxel_downsample(kitti[1], 0.1)[0]
    idx, d2, _ = impl.knn_covariances(f1, 20)
    k = pcl["knn"]["cloud1@0.1,k=20"]
    report["kNN-20 indices"] = fnv(idx.astype(np.uint32)) == k["index_fnv1a"]
    report["kNN-20 distances"] = fnv(d2.astype(np.float32)) == k["dist2_fnv1a"]
    nn_i, nn_d = impl.nearest_neighbors(f1, f2)
    k = pcl["knn"]["nn cloud2@0.1 -> cloud1@0.1"]
    report["1-NN indices"] = fnv(nn_i.astype(np.uint32)) == k["index_fnv1a"]
    report["1-NN distances"] = fnv(nn_d.astype(np.float32)) == k["dist2_fnv1a"]
    return report


def _optimizer_of(pcl):
    if "inner_optimizer" in pcl:
        return pcl["inner_optimizer"]
    major, minor = (int(v) for v in pcl["pcl_version"].split(".")[:2])
    return "newton" if (major, minor) >= (1, 14) else "bfgs"


def _check_align(pcl, impl, kitti, tol=(1e-4, 1e-4, 1e-4)):
    worst = (0.0, 0.0, 0.0)
    for key, g in pcl["align"].items():
        pair, density = key.split("@")
        a, b = (int(s[-1]) - 1 for s in pair.split("->"))
        r = impl.gicp_align(kitti[a], kitti[b], None, RegistrationParameters.defaults(point_cloud_density=float(density)))
        assert r.status == g["status"] and r.converged == g["converged"], key
        assert (r.n_source, r.n_target) == (g["n_source"], g["n_target"]), key
        dt, dr = pose_delta(np.array(g["T"]), r.pose())
        df = abs(r.fitness - g["fitness"]) / max(abs(g["fitness"]), 1e-12)
        assert dt < tol[0] and dr < tol[1] and df < tol[2], (key, dt, dr, df)
        worst = (max(worst[0], dt), max(worst[1], dr), max(worst[2], df))
    return worst


@pytest.mark.parametrize("path", cases())
def test_oracle_matches_pcl(path, oracle_mod, kitti):
    if path is None:
        pytest.skip(SKIP)
    pcl = json.load(open(path))
    report = _check_voxel_knn(pcl, oracle_mod, kitti, oracle_mod.fnv1a)
    assert all(report.values()), {k: v for k, v in report.items() if not v}
    old = oracle_mod.set_gicp_optimizer(_optimizer_of(pcl))
    try:
        print("PCL", pcl["pcl_version"], _optimizer_of(pcl), "vs oracle: worst |dt|, |dr|, fitness rel.:", _check_align(pcl, oracle_mod, kitti))
    finally:
        oracle_mod.set_gicp_optimizer(old)


@pytest.mark.gpu
@pytest.mark.parametrize("path", cases())
def test_gpu_matches_pcl(path, oracle_mod, kitti):
    if path is None:
        pytest.skip(SKIP)
    import slam3d_b200
    pcl = json.load(open(path))
    ctx = slam3d_b200.Context()
    report = _check_voxel_knn(pcl, ctx, kitti, oracle_mod.fnv1a)
    assert all(report.values()), {k: v for k, v in report.items() if not v}
    tol = (1e-4, 1e-4, 1e-4) if _optimizer_of(pcl) == "newton" else (1e-2, 1e-3, 1e-2)
    print("PCL", pcl["pcl_version"], _optimizer_of(pcl), "vs GPU: worst |dt|, |dr|, fitness rel.:", _check_align(pcl, ctx, kitti, tol))
    ctx.close()


def test_driver_and_cmake_are_in_place():
    """The ready-to-run pieces exist and name what they must: the swap, the setter sequence, the version floor."""
    root = os.path.dirname(HERE)
    src = open(os.path.join(root, "baseline", "doicp_driver.cpp")).read()
    for needle in ("GeneralizedIterativeClosestPoint", "setInputSource(ft)", "setInputTarget(fs)", "getFitnessScore(cfg.max_correspondence_distance)",
                   "setCorrespondenceRandomness", "pcl::VoxelGrid", "guess.matrix().cast<float>()"):
        assert needle in src, needle
    cm = open(os.path.join(root, "baseline", "CMakeLists.txt")).read()
    assert "find_package(PCL 1.8.1" in cm


def test_consumer_runs_on_a_self_made_golden(oracle_mod, kitti):
    """Keeps the consumer above from rotting while no PCL file exists: a file in the driver's format is made from the oracle
    itself (BFGS mode, as a PCL <= 1.13 would be) and pushed through the same checks — it must pass in BFGS mode and be told
    apart from the Newton mode at the strict tolerance."""
    fnv = oracle_mod.fnv1a
    pcl = {"pcl_version": "1.12.1", "voxel": {}, "knn": {}, "align": {}}
    assert _optimizer_of(pcl) == "bfgs" and _optimizer_of({"pcl_version": "1.14.0"}) == "newton"
    for leaf in ("0.5", "1"):
        pcl["voxel"][leaf] = []
        for c in kitti:
            out = oracle_mod.voxel_downsample(c, float(leaf))[0]
            pcl["voxel"][leaf].append({"n_out": int(out.shape[0]), "xyz_fnv1a": fnv(np.ascontiguousarray(out[:, :3]))})
    f1 = oracle_mod.voxel_downsample(kitti[0], 0.1)[0]
    f2 = oracle_mod.voxel_downsample(kitti[1], 0.1)[0]
    idx, d2, _ = oracle_mod.knn_covariances(f1, 20)
    pcl["knn"]["cloud1@0.1,k=20"] = {"index_fnv1a": fnv(idx.astype(np.uint32)), "dist2_fnv1a": fnv(d2.astype(np.float32))}
    nn_i, nn_d = oracle_mod.nearest_neighbors(f1, f2)
    pcl["knn"]["nn cloud2@0.1 -> cloud1@0.1"] = {"index_fnv1a": fnv(nn_i.astype(np.uint32)), "dist2_fnv1a": fnv(nn_d.astype(np.float32))}
    old = oracle_mod.set_gicp_optimizer("bfgs")
    try:
        r = oracle_mod.gicp_align(kitti[0], kitti[1], None, RegistrationParameters.defaults(point_cloud_density=0.2))
        pcl["align"]["cloud1->cloud2@0.2"] = {"status": r.status, "converged": r.converged, "n_source": r.n_source, "n_target": r.n_target,
                                              "T": np.asarray(r.pose()).tolist(), "fitness": r.fitness}
        report = _check_voxel_knn(pcl, oracle_mod, kitti, fnv)
        assert all(report.values()), report
        assert max(_check_align(pcl, oracle_mod, kitti)) < 1e-12  # the same run twice (pose_delta itself rounds at 1e-16)
    finally:
        oracle_mod.set_gicp_optimizer(old)
    with pytest.raises(AssertionError):  # Newton mode against a BFGS file: millimetres apart, refused at 1e-4
        _check_align(pcl, oracle_mod, kitti)
    _check_align(pcl, oracle_mod, kitti, tol=(1e-2, 1e-3, 1e-2))  # ... and inside what the CUDA path is held to against such a file
